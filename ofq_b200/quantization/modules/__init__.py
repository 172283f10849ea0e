from .attention import QAttention, QAttention_qkreparam, QAttention_qkreparam_4_cga
from .qbias import LearnableBias, LearnableBias4img
from .qlinear import LSQ_input, LSQ_QConv2d, LSQ_QLinear4head, QLinear, QMLP
from .swin_attention_and_mlp import (QAttention_swin, QAttention_swin_qkreparam, QAttention_swin_qkreparam_4_cga,
                                     QMLP_swin)
from .utils import (deit_qmodule_names, get_module_by_name, make_qconfigs, replace_module_by_qmodule_deit,
                    replace_module_by_qmodule_swin, set_module_by_name, swin_qmodule_names)
