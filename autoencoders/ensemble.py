from sparse_coding_b200.ensemble import (FunctionalEnsemble, construct_stacked_leaf, optim_str_to_func,  # noqa: F401
                                         stack_dict, unstack_dict)
from sparse_coding_b200.signatures import DictSignature  # noqa: F401
