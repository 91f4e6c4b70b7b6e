"""
FITS staging: Observation.save_observation / save_mapped_observation
(planetmapper/observation.py:1185-1474).

CPU tests pin the header-card formatter and HDU layout against the raw cards of the golden
files the reference's own tests compare (tests/test_observation.py:1016-1280; cards exported
to tests/golden/ref_cards.json by tests/golden/make_golden.py).  GPU tests check the staged
file image bit-for-bit against oracle/fits_oracle.py and the saved files against the golden
arrays.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from fits_min import read_cards, read_fits
from helpers import PLANE_NAMES, max_diff
from oracle import fits_oracle as FO
from planetmapper_b200 import fits_stage as FS
from test_oracle_golden import GOLDEN_TOL, MAP_FILES, WRAP


@pytest.fixture(scope='module')
def ref_cards():
    with open(os.path.join(GOLDEN, 'ref_cards.json')) as f:
        return json.load(f)


def _observation(bc_hst, golden_arrays, golden_headers):
    import planetmapper_b200 as pm

    hdr = golden_headers['inputs/test.fits']
    user = [(k, hdr[k], None) for k in ('TARGET', 'TELESCOP', 'DATE-OBS', 'TIME-OBS', 'CUSTOM')]
    o = pm.Observation(data=golden_arrays['inputs/test.fits/PRIMARY'], constants=bc_hst,
                       header=FS.Header(user))
    o.set_disc_params(2.5, 3.1, 3.9, 123.456)
    o.set_disc_method('<<<test>>>')
    return o


# ---- CPU: card text and layout ---------------------------------------------------------------
def test_every_golden_card_is_reproduced_exactly(ref_cards):
    n = 0
    for fn, hdus in ref_cards['cards'].items():
        for cards in hdus:
            for card in cards:
                key, value, comment = FO.parse_card(card)
                got = FS.Header.format_card(key, value, comment)
                assert got == [card], f'{fn}: {card!r} -> {got!r}'
                n += 1
    assert n > 300


def test_value_formatting_rules():
    f = FS.Header.format_card
    assert f('NAXIS1', 7, None)[0].rstrip() == 'NAXIS1  =                    7'
    assert f('CRVAL1', 345.0, None)[0].rstrip() == 'CRVAL1  =                345.0'
    assert f('X', 1e-30, None)[0].rstrip() == 'X       =                1E-30'
    assert f('X', -1.2345678901234567e+100, None)[0].rstrip() == 'X       = -1.234567890123E+100'
    assert f('NAME', "it's", None)[0].rstrip() == "NAME    = 'it''s   '"
    assert f('EMPTY', '', None)[0].rstrip() == "EMPTY   = ''"
    assert f('COMMENT', 'x' * 100, None) == [('COMMENT ' + 'x' * 72), ('COMMENT ' + 'x' * 28).ljust(80)]
    assert f('PLANMAP DISC X0', 2.5, 'c')[0].rstrip() == 'HIERARCH PLANMAP DISC X0 = 2.5 / c'
    assert all(len(c) == 80 for c in f('PLANMAP A VERY LONG KEYWORD', 1.5, 'comment ' * 20))
    with pytest.raises(FS.VerifyError):
        f('X', float('nan'), None)
    with pytest.raises(FS.VerifyError):
        f('PLANMAP ' + 'K' * 70, 'value', None)
    with pytest.raises(FS.VerifyError):
        f('BAD*KEY', 1, None)


def test_card_format_round_trip_property():
    """Property (hypothesis): every value the formatter accepts parses back to itself - floats
    bit-exactly whenever their shortest repr fits FITS's 20-character value field - and every card
    is exactly 80 ASCII characters."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    keys = st.sampled_from(['NAXIS1', 'CRVAL1', 'PLANMAP DISC X0', 'PLANMAP MAP SMOOTH-MAX-OVERSAMPLED-IMG-SIZE', 'A'])
    text = st.text(alphabet=st.characters(min_codepoint=32, max_codepoint=126), max_size=18)
    values = st.one_of(st.booleans(), st.integers(min_value=-10**15, max_value=10**15),
                       st.floats(allow_nan=False, allow_infinity=False), text)

    @settings(max_examples=400, deadline=None, derandomize=True)
    @given(keys, values, st.one_of(st.none(), text.filter(lambda t: t.strip() == t and t != '' and '/' not in t)))
    def check(key, value, comment):
        cards = FS.Header.format_card(key, value, comment)
        assert len(cards) == 1 and len(cards[0]) == 80 and cards[0].isascii()
        k, v, c = FO.parse_card(cards[0])
        assert k == key
        if isinstance(value, str):
            assert v == value.rstrip()
        elif isinstance(value, float):
            assert isinstance(v, (float, int))
            if len(str(value)) <= 20:
                assert float(v) == value
            else:
                assert float(v) == pytest.approx(value, rel=1e-11)   # astropy keeps 20 characters: >= 12 digits
        else:
            assert v == value and type(v) is type(value)
        if isinstance(value, str):
            # the card keeps trailing blanks inside the quotes (as astropy does) while parsing strips
            # them, so the re-format identity holds for the stripped value
            assert FS.Header.format_card(k, v, c) == FS.Header.format_card(key, value.rstrip(), comment)
        elif not isinstance(value, float):
            assert FS.Header.format_card(k, v, c) == cards

    check()


def test_header_object_semantics():
    h = FS.Header([('A', 1, None)])
    h['B'] = 2.0
    h['A'] = 3
    h.append('HIERARCH PLANMAP X', 'v', 'c')
    assert h.keys() == ['A', 'B', 'PLANMAP X'] and h['A'] == 3 and 'planmap x' in h
    h.remove('B')
    h.remove('B', ignore_missing=True)
    with pytest.raises(KeyError):
        h.remove('B', ignore_missing=False)
    c = h.copy()
    c['A'] = 4
    assert h['A'] == 3 and len(c) == 2


def test_extension_header_blocks_match_golden(ref_cards):
    for fn, shape in (('test_nav.fits', (10, 7)), ('map_rectangular-linear.fits', (6, 12)),
                      ('map_orthographic-1.fits', (10, 10))):
        for cards in ref_cards['cards'][fn][1:]:
            parsed = [FO.parse_card(c) for c in cards]
            kv = {k: v for k, v, _ in parsed}
            user = FS.Header([p for p in parsed if p[0] in ('ABOUT', 'COMMENT')
                              or p[0][:5] in ('CTYPE', 'CUNIT', 'CRPIX', 'CRVAL', 'CDELT')])
            blocks = FS.hdu_header_bytes(shape, user, primary=False, name=kv['EXTNAME'])
            want = ''.join(cards) + 'END'.ljust(80)
            want += ' ' * (-len(want) % 2880)
            assert blocks.decode() == want, f'{fn} {kv["EXTNAME"]}'


def test_primary_header_of_save_observation_matches_golden(bc_hst, golden_arrays, golden_headers, ref_cards):
    obs = _observation(bc_hst, golden_arrays, golden_headers)
    header = obs.header.copy()
    obs.add_header_metadata(header)
    blocks = FS.hdu_header_bytes(obs.data.shape, header, primary=True)
    assert len(blocks) % 2880 == 0
    got = [blocks[i:i + 80].decode() for i in range(0, len(blocks), 80)]
    got = got[:got.index('END'.ljust(80))]
    ref = [c for c in ref_cards['cards']['test_nav.fits'][0] if 'PLANMAP INFILE' not in c]  # no input file here
    assert len(got) == len(ref)
    free = ('PLANMAP VERSION', 'PLANMAP DATE')
    for g, r in zip(got, ref):
        gk, gv, gc = FO.parse_card(g)
        rk, rv, rc = FO.parse_card(r)
        assert gk == rk, (g, r)
        if gk in free:
            continue
        if isinstance(rv, float):
            assert gv == pytest.approx(rv, rel=1e-11, abs=1e-9), gk
            if len(g.rstrip()) < 80 and len(r.rstrip()) < 80 and str(gv) == str(rv):
                assert g == r
        else:
            assert (gv, gc) == (rv, rc), gk
            assert g == r


@pytest.mark.parametrize('fn', ['map_rectangular-linear.fits', 'map_rectangular-smooth.fits',
                                'map_rectangular-cubic.fits', 'map_orthographic-1.fits', 'map_azimuthal-1.fits'])
def test_map_metadata_cards_match_golden(bc_hst, golden_arrays, golden_headers, ref_cards, fn):
    """_add_map_header_metadata / _add_map_wcs_to_header (observation.py:1476-1612).  Only the
    rectangular projections run here: the others need the device for generate_map_coordinates."""
    import torch

    spec, alt, interp = MAP_FILES[fn]
    if spec[0] != 'rectangular' and not torch.cuda.is_available():
        pytest.skip('projection grids are computed on the device')
    obs = _observation(bc_hst, golden_arrays, golden_headers)
    kw = dict(degree_interval=spec[1]) if spec[0] == 'rectangular' else dict(projection=spec[0], lon=spec[1],
                                                                              lat=spec[2], size=spec[3])
    header = FS.Header()
    obs._add_map_header_metadata(header, interpolation=interp, spline_smoothing=0, propagate_nan=True,
                                 smooth_oversample_by=5, smooth_max_oversampled_img_size=10_000, **kw)
    obs._add_map_wcs_to_header(header, **kw)
    got = header.card_images()
    ref = [c for c in ref_cards['cards'][fn][0] if 'PLANMAP MAP' in c or c[:5] in ('CTYPE', 'CUNIT', 'CRPIX', 'CRVAL',
                                                                                  'CDELT')]
    assert got == ref


def test_backplane_descriptions_match_golden_about_cards(bc_hst, ref_cards):
    import planetmapper_b200 as pm

    body = pm.BodyXY(constants=bc_hst, nx=7, ny=10)
    about = {k: v for k, v in ref_cards['about'].items() if k != 'WIREFRAME'}
    assert {n: bp.description for n, bp in body.backplanes.items()} == about


def test_map_function_params_are_consistent(bc_hst, golden_arrays, golden_headers):
    """tests/test_observation.py:815-860: map_img, get_mapped_data, save_mapped_observation and
    _add_map_header_metadata share parameter names, annotations and defaults."""
    import inspect

    obs = _observation(bc_hst, golden_arrays, golden_headers)

    def compare(p1, p2, allow_empty_default=False):
        assert p1.name == p2.name and p1.annotation == p2.annotation, p1.name
        empty = inspect.Parameter.empty
        if not allow_empty_default or (p1.default is not empty and p2.default is not empty):
            assert p1.default == p2.default, p1.name

    map_img = inspect.signature(obs.map_img).parameters
    func = inspect.signature(obs.get_mapped_data).parameters
    for k in set(func) | (set(map_img) - {'img', 'warn_nan'}):
        compare(func[k], map_img[k])
    func = inspect.signature(obs.save_mapped_observation).parameters
    for k in set(map_img) - {'img', 'warn_nan'}:
        compare(func[k], map_img[k])
    func = inspect.signature(obs._add_map_header_metadata).parameters
    for k in set(map_img) - {'img', 'warn_nan'}:
        compare(func[k], map_img[k], allow_empty_default=True)


def test_append_to_header_make_filename_and_names_to_save(bc_hst):
    """tests/test_observation.py:862-1014."""
    import planetmapper_b200 as pm

    obs = pm.Observation(data=np.ones((5, 10, 8)), constants=bc_hst)
    obs.append_to_header('TESTING', 123, 'Testing comment')
    assert obs.header['HIERARCH PLANMAP TESTING'] == 123
    assert obs.header.comments['HIERARCH PLANMAP TESTING'] == 'Testing comment'
    header = FS.Header()
    obs.append_to_header('TESTING', 123, 'Testing comment', header=header)
    assert header['HIERARCH PLANMAP TESTING'] == 123 and 'TESTING' not in header
    header = FS.Header()
    obs.append_to_header('TESTING', 123, 'Testing comment', header=header, hierarch_keyword=False)
    assert header['TESTING'] == 123 and header.comments['TESTING'] == 'Testing comment'
    assert 'HIERARCH PLANMAP TESTING' not in header
    header = FS.Header()
    obs.append_to_header('A', 0, header=header, hierarch_keyword=False)
    obs.append_to_header('B', 1, header=header, hierarch_keyword=False)
    obs.append_to_header('A', 1, header=header, hierarch_keyword=False)
    assert header['A'] == 1 and header.keys() == ['B', 'A']
    header = FS.Header()
    obs.append_to_header('A', 0, header=header, hierarch_keyword=False)
    obs.append_to_header('B', 1, header=header, hierarch_keyword=False)
    obs.append_to_header('A', 1, header=header, hierarch_keyword=False, remove_existing=False)
    assert header['A'] == 0 and header.keys() == ['A', 'B', 'A']
    for n in range(100):
        s = 'x' * n
        obs.append_to_header('TESTING', s)
        if n >= 53:
            s = 'x' * 49 + '...'
        assert obs.header['HIERARCH PLANMAP TESTING'] == s
        assert all(len(c) == 80 for c in obs.header.card_images())
    obs.append_to_header('TESTING', 'x' * 100, truncate_strings=False)
    assert obs.header['HIERARCH PLANMAP TESTING'] == 'x' * 100
    with pytest.raises(ValueError):   # astropy < 7.1 behaviour: no CONTINUE cards for HIERARCH keywords
        obs.header.card_images()

    obs = pm.Observation(data=np.ones((5, 10, 8)), constants=bc_hst)
    obs.add_header_metadata()
    assert 'HIERARCH PLANMAP INFILE' not in obs.header and obs.header['PLANMAP TARGET'] == 'JUPITER'
    assert obs.make_filename() == 'JUPITER_2005-01-01T000000.fits'
    assert obs.make_filename(extension='.txt') == 'JUPITER_2005-01-01T000000.txt'
    assert obs.make_filename(prefix='pre_', suffix='_post') == 'pre_JUPITER_2005-01-01T000000_post.fits'

    assert obs._get_backplane_names_to_save(None, frozenset()) == set(PLANE_NAMES)
    assert obs._get_backplane_names_to_save(['RA', 'DEC'], frozenset()) == {'RA', 'DEC'}
    assert obs._get_backplane_names_to_save(['RA', 'DEC'], ['RA']) == {'DEC'}
    assert obs._get_backplane_names_to_save(
        backplanes_to_save=['RA', '   dec   ', 'DISTANCE', 'radial-VELOCITY', '<some other backplane>'],
        backplanes_to_skip=['DEC', 'dISTANCE   ', 'LIMB-DISTANCE']) == {'RA', 'RADIAL-VELOCITY', '<SOME OTHER BACKPLANE>'}


def test_file_layout_offsets():
    hdus = [FS.ImageHDU(np.zeros((10, 10, 7))), FS.ImageHDU(np.zeros((10, 7)), name='A'),
            FS.ImageHDU(np.zeros((0, 7)), name='EMPTY'), FS.ImageHDU(np.zeros((360, 1)), name='B')]
    headers, hoff, doff, size = FS.file_layout(hdus)
    assert [len(h) for h in headers] == [2880] * 4
    assert hoff == [0, 2880 + 5760, 2880 + 5760 + 2880 + 2880, 2880 + 5760 + 2880 + 2880 + 2880]
    assert doff == [o + 2880 for o in hoff]
    assert size == doff[-1] + 2880


def test_abi_argument_checks():
    from planetmapper_b200 import _lib as L

    lib = L.load_library()
    assert lib.pm_fits_data_unit_bytes(0) == 0
    assert lib.pm_fits_data_unit_bytes(70) == 2880
    assert lib.pm_fits_data_unit_bytes(360) == 2880
    assert lib.pm_fits_data_unit_bytes(361) == 5760
    assert lib.pm_fits_data_unit_bytes(-1) == -1
    assert lib.pm_fits_stage(None, None, None, 1, None, None) == -1
    assert lib.pm_fits_stage(None, None, None, 0, None, None) == 0


# ---- GPU: staged bytes and saved files -------------------------------------------------------
@pytest.mark.gpu
def test_staged_image_is_bit_exact_against_the_oracle():
    import torch

    rng = np.random.default_rng(5)
    shapes = [(3, 5, 7), (1,), (0, 4), (360,), (361,), (719, 3)] + [(int(rng.integers(1, 2000)),) for _ in range(40)]
    arrays = []
    for s in shapes:
        a = rng.normal(size=s) * 10.0 ** rng.integers(-300, 300, size=s)
        flat = a.reshape(-1)
        if flat.size:
            special = np.array([np.nan, np.inf, -np.inf, -0.0, 5e-324, 1.7976931348623157e308])
            idx = rng.integers(0, flat.size, size=min(6, flat.size))
            flat[idx] = special[:idx.size]
        arrays.append(a)
    hdus = [FS.ImageHDU(a, name=None if i == 0 else f'E{i}') for i, a in enumerate(arrays)]
    assert len(hdus) > 32  # more than one launch
    host = FS.stage_file_image(hdus)
    headers, *_ = FS.file_layout(hdus)
    assert host.numpy().tobytes() == FO.assemble(headers, arrays)
    # a large unit, device-resident input, non-zero destination garbage
    big = torch.randn(3_000_001, dtype=torch.float64, device='cuda')
    host = FS.stage_file_image([FS.ImageHDU(big)])
    headers, *_ = FS.file_layout([FS.ImageHDU(big)])
    assert host.numpy().tobytes() == FO.assemble(headers, [big.cpu().numpy()])


def _check(a, ref, name, label):
    assert a.shape == ref.shape and a.dtype == np.float64
    assert np.array_equal(np.isnan(a), np.isnan(ref)), f'{label} {name}: NaN mask'
    d = max_diff(a, ref, wrap=name in WRAP)
    assert d <= GOLDEN_TOL[name], f'{label} {name}: {d:.3e}'


@pytest.mark.gpu
def test_save_observation_file_matches_golden(tmp_path, bc_hst, golden_arrays, golden_headers, ref_cards):
    from planetmapper_b200 import _lib as L

    obs = _observation(bc_hst, golden_arrays, golden_headers)
    path = tmp_path / 'sub' / 'nav.fits'
    before = L.launch_count()
    obs.save_observation(path, print_info=False)
    assert L.launch_count() - before == 2  # one fused backplane launch + one staging launch
    hdus = read_fits(path)
    want = [n for n in ref_cards['extnames']['test_nav.fits'] if n != 'WIREFRAME']
    assert [h.get('EXTNAME', 'PRIMARY') for h, _ in hdus] == want == ['PRIMARY'] + PLANE_NAMES
    assert os.path.getsize(path) % 2880 == 0
    assert np.array_equal(hdus[0][1], golden_arrays['inputs/test.fits/PRIMARY'], equal_nan=True)
    for hdr, arr in hdus[1:]:
        _check(arr, golden_arrays[f'test_nav.fits/{hdr["EXTNAME"]}'], hdr['EXTNAME'], 'saved nav')
        assert hdr['BITPIX'] == -64 and (hdr['NAXIS1'], hdr['NAXIS2']) == (7, 10)
    # the saved planes are exactly what the getters return
    for hdr, arr in hdus[1:]:
        assert np.array_equal(arr, obs.get_backplane_img(hdr['EXTNAME']), equal_nan=True)
    # extension headers are byte-identical to the reference's
    cards = read_cards(path)
    assert cards[1] == ref_cards['cards']['test_nav.fits'][1]
    assert cards[2] == ref_cards['cards']['test_nav.fits'][2]

    # selection, altitude and a user-registered backplane
    obs.register_backplane('CUSTOM', 'custom plane [1]', lambda: np.full((10, 7), 4.25), lambda **kw: None)
    obs.save_observation(path, backplanes_to_save=['ra', 'DEC', 'Distance', 'custom'], backplanes_to_skip=['dec'],
                         print_info=False, alt=34567.8912)
    hdus = read_fits(path)
    assert [h.get('EXTNAME', 'PRIMARY') for h, _ in hdus] == ['PRIMARY', 'RA', 'DISTANCE', 'CUSTOM']
    assert hdus[0][0]['PLANMAP ALTITUDE-ADJUSTMENT'] == 34567.8912
    _check(hdus[2][1], golden_arrays['test_nav_alt.fits/DISTANCE'], 'DISTANCE', 'saved nav alt')
    assert np.array_equal(hdus[3][1], np.full((10, 7), 4.25))
    assert obs._alt_adjustment == 0.0
    with pytest.raises(NotImplementedError):
        obs.save_observation(path, include_wireframe=True, print_info=False)


@pytest.mark.gpu
@pytest.mark.parametrize('fn', ['map_rectangular-linear.fits', 'map_rectangular-nearest.fits',
                                'map_rectangular-smooth.fits', 'map_rectangular-cubic.fits',
                                'map_orthographic-1.fits', 'map_azimuthal-1.fits'])
def test_save_mapped_observation_file_matches_golden(tmp_path, bc_hst, golden_arrays, golden_headers, ref_cards, fn):
    spec, alt, interp = MAP_FILES[fn]
    kw = dict(degree_interval=spec[1]) if spec[0] == 'rectangular' else dict(projection=spec[0], lon=spec[1],
                                                                              lat=spec[2], size=spec[3])
    obs = _observation(bc_hst, golden_arrays, golden_headers)
    path = tmp_path / fn
    have_planes = f'{fn}/LON-GRAPHIC' in golden_arrays
    obs.save_mapped_observation(path, interpolation=interp, include_backplanes=have_planes, print_info=False, **kw)
    hdus = read_fits(path)
    names = [h.get('EXTNAME', 'PRIMARY') for h, _ in hdus]
    assert names == ['PRIMARY'] + (PLANE_NAMES if have_planes else [])
    mapped, ref = hdus[0][1], golden_arrays[f'{fn}/PRIMARY']
    assert mapped.shape == ref.shape
    assert np.array_equal(np.isnan(mapped), np.isnan(ref))
    ok = np.isfinite(ref)
    if interp == 'nearest':
        assert np.array_equal(mapped[ok], ref[ok])
    elif ok.any():
        assert np.max(np.abs(mapped[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1.0)) < 1e-8
    assert np.array_equal(mapped, obs.get_mapped_data(interp, **kw), equal_nan=True)
    for hdr, arr in hdus[1:]:
        _check(arr, golden_arrays[f'{fn}/{hdr["EXTNAME"]}'], hdr['EXTNAME'], fn)
    if fn in ref_cards['cards']:
        cards = read_cards(path)
        ref_primary = ref_cards['cards'][fn][0]
        keep = lambda c: 'PLANMAP MAP' in c or not c.startswith('HIERARCH')  # noqa: E731
        assert [c for c in cards[0] if keep(c)] == [c for c in ref_primary if keep(c)]
        if have_planes:
            assert cards[1] == ref_cards['cards'][fn][1]
