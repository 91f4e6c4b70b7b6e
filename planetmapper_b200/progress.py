"""
Progress hooks (the reference's planetmapper/progress.py:16-41 and base.py:760-781): an object may carry a
callable ``hook(progress, stack)``; decorated methods report 0 when they start and 1 when they finish, with
``stack`` the qualified names of the decorated calls currently running, and long loops report fractions in
between.  On this path a "long loop" is a handful of kernel launches, so the useful updates come from the
chunked cube mapping and the save methods; the mechanism is the reference's so that GUI-style consumers
(cancel by raising from the hook, progress bars) keep working.
"""
from __future__ import annotations

from functools import wraps
from typing import Callable


class ProgressHook:
    """Base class: ``__call__(progress in [0, 1], stack of qualified names)``."""

    def __call__(self, progress: float, stack: list[str]) -> None:   # pragma: no cover - interface
        raise NotImplementedError


class CLIProgressHook(ProgressHook):
    """One tqdm bar per running decorated call (closed when it reports 1)."""

    def __init__(self, leave: bool | None = None) -> None:
        self.leave = leave
        self.bars: dict = {}

    def __call__(self, progress: float, stack: list[str]) -> None:
        import tqdm

        key = tuple(stack)
        bar = self.bars.get(key)
        if bar is None:
            bar = self.bars[key] = tqdm.tqdm(total=100, desc=stack[-1] if stack else '', unit='%',
                                             leave=(len(stack) == 1) if self.leave is None else self.leave)
        bar.update(progress * 100 - bar.n)
        if progress == 1:
            bar.close()
            self.bars.pop(key, None)


class ProgressMixin:
    """``_set_progress_hook`` / ``_get_progress_hook`` / ``_remove_progress_hook`` / ``_update_progress_hook``
    with the reference's semantics (base.py:760-781): the hook may raise to cancel, so updates are only sent
    from places where an exception leaves no side effect."""

    _progress_hook: Callable | None = None
    _progress_call_stack: list

    def _set_progress_hook(self, progress_hook) -> None:
        self._progress_hook = progress_hook
        self._progress_call_stack = []

    def _get_progress_hook(self):
        return self._progress_hook

    def _remove_progress_hook(self) -> None:
        self._progress_hook = None
        self._progress_call_stack = []

    def _update_progress_hook(self, progress_frac: float) -> None:
        if self._progress_hook is not None:
            self._progress_hook(progress_frac, self._progress_call_stack)


def progress_decorator(fn):
    """Report 0 / 1 around ``fn`` and keep the call stack (progress.py:16-41)."""

    @wraps(fn)
    def decorated(self, *args, **kwargs):
        if getattr(self, '_progress_hook', None) is None:
            return fn(self, *args, **kwargs)
        self._progress_call_stack.append(fn.__qualname__)
        try:
            self._update_progress_hook(0)     # inside the try: a hook that cancels here must not leave its frame behind
            out = fn(self, *args, **kwargs)
        except BaseException:
            self._progress_call_stack.pop()
            raise
        self._update_progress_hook(1)
        if self._progress_call_stack:
            self._progress_call_stack.pop()
        return out

    return decorated
