"""Drop-in module name of the reference (`from utils import *`)."""
from dorpatch_b200.utils import *  # noqa: F401,F403
from dorpatch_b200.utils import (NUM_CLASSES_DICT, NormModel, clip, convert_float_list_to_str,  # noqa: F401
                                 generate_saving_path, get_dataset, get_model, get_normalize, set_device,
                                 set_random_seed)
