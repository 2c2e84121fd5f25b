"""No-op stand-in for `colorama` (absent in this image) so the reference's display
modules import. Test infrastructure only: used by oracle/gen_golden.py."""


class _Codes:
    def __getattr__(self, name):
        return ''


Style = _Codes()
Fore = _Codes()
Back = _Codes()


def init(*args, **kwargs):
    pass
