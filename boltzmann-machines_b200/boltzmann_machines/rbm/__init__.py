from .base_rbm import *
from .rbm import *
