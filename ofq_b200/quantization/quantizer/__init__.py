from .lsq import LsqQuantizer, LsqQuantizer4v, LsqQuantizerWeight
from .statsq import StatsQuantizer, StatsQuantizer_specific_4_qkreparam_cga
