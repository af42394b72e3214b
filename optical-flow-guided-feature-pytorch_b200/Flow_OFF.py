"""Drop-in for the reference's ``Flow_OFF.py``: same class and factory names (``from off_b200.Flow_OFF import bninception_off``),
OFF section on the B200 path.  See off_b200/models.py."""

from . import models as _m


class BNInception_OFF(_m.BNInception_OFF):
    def __init__(self, num_classes=1000, batch=16, length=7, **kw):
        kw.setdefault("variant", "flow")
        super().__init__(num_classes, batch, length, **kw)


def bninception_off(num_classes=1000, batch=1, num_seg=25, **kw):
    return BNInception_OFF(num_classes, batch, num_seg, **kw)
