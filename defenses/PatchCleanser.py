"""Drop-in module path of the reference (pickled PatchCleanserRecord objects resolve here)."""
from dorpatch_b200.defenses.PatchCleanser import (MaskWindow, PatchCleanser, PatchCleanserRecord,  # noqa: F401
                                                  PatchCleanserResult)
