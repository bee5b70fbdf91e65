from .abstract_learner import Learner
from .backpropagation import Backpropagation
from .neuroevolution import NeuroEvolution
