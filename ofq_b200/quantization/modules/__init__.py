from .attention import QAttention, QAttention_qkreparam, QAttention_qkreparam_4_cga
from .qbias import LearnableBias, LearnableBias4img
from .qlinear import LSQ_input, LSQ_QConv2d, LSQ_QLinear4head, QLinear, QMLP
from .utils import (deit_qmodule_names, get_module_by_name, make_qconfigs, replace_module_by_qmodule_deit,
                    set_module_by_name)
