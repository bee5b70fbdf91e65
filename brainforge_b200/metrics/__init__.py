from .costs import (mean_squared_error, categorical_crossentropy, binary_crossentropy, hinge, mse, cxent, bxent)
from .metrics import accuracy
from . import costs, metrics
