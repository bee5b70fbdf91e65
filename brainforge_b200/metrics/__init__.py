from .costs import mean_squared_error, categorical_crossentropy, mse, cxent
from .metrics import accuracy
from . import costs, metrics
