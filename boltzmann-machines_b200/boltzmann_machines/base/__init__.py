from .base import *
from .base_model import *
from .mixin import *
from .native_model import *
from . import tf_model
