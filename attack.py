"""Drop-in module name of the reference (`from attack import DorPatch`)."""
from dorpatch_b200.attack import (CW_loss, DorPatch, get_mask_set, local_variance,  # noqa: F401
                                  min_var_weighted_variance)
