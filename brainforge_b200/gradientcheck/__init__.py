from .gradientcheck import run
from .raw_gradients import analytical_gradients, numerical_gradients
