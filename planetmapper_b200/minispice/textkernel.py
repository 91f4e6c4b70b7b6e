"""
Parser for NAIF text kernels (PCK ``.tpc``, LSK ``.tls``): only the data blocks
between ``\\begindata`` and ``\\begintext`` are read. Follows the public "Kernel
Required Reading" syntax: ``NAME = value`` or ``NAME = ( v1 v2 ... )``, ``+=``
appends, Fortran ``D`` exponents, quoted strings and ``@date`` tokens.

Host-side, once per process. The reference reaches the same values through
``spice.bodvar`` (planetmapper/body.py:522, :528).
"""

from __future__ import annotations

import re

_TOKEN = re.compile(r"'(?:[^']|'')*'|[^\s,()]+|[()]")


def _convert(tok: str):
    if tok.startswith("'"):
        return tok[1:-1].replace("''", "'")
    if tok.startswith('@'):
        return tok
    try:
        return float(tok.replace('D', 'E').replace('d', 'e'))
    except ValueError:
        return tok


def parse_text_kernel(text: str) -> dict[str, list]:
    pool: dict[str, list] = {}
    in_data = False
    data_lines: list[str] = []
    for line in text.splitlines():
        stripped = line.strip()
        if stripped.startswith('\\begindata'):
            in_data = True
            continue
        if stripped.startswith('\\begintext'):
            in_data = False
            continue
        if in_data:
            data_lines.append(line)
    tokens = _TOKEN.findall('\n'.join(data_lines))
    i = 0
    n = len(tokens)
    while i < n:
        name = tokens[i]
        if i + 1 >= n or tokens[i + 1] not in ('=', '+='):
            i += 1
            continue
        op = tokens[i + 1]
        i += 2
        values = []
        if i < n and tokens[i] == '(':
            i += 1
            while i < n and tokens[i] != ')':
                values.append(_convert(tokens[i]))
                i += 1
            i += 1
        elif i < n:
            values.append(_convert(tokens[i]))
            i += 1
        if op == '+=' and name in pool:
            pool[name].extend(values)
        else:
            pool[name] = values
    return pool


def load_text_kernel(path: str) -> dict[str, list]:
    with open(path, 'r', encoding='latin-1') as f:
        return parse_text_kernel(f.read())
