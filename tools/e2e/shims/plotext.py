"""No-op stand-in for `plotext` (terminal plotting; absent from this image, no network).

The reference's CLI script only uses it to draw the PSNR curve after training; every attribute is a function
that accepts anything and does nothing.
"""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None

    return _noop
