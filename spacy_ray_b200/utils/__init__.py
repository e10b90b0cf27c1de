from .logging import logger
from .code import import_code
from .timing import CudaStepTimer, ClockSampler

__all__ = ["logger", "import_code", "CudaStepTimer", "ClockSampler"]
