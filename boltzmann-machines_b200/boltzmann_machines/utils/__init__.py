from .rng import *
from .utils import *
from .stopwatch import *
from . import testing
