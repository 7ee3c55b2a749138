from .modules import ActFun, ConvLSTM, ConvLSTMLayer  # noqa: F401
from .utils import batch_reduce, set_gpu, split_feature  # noqa: F401
