from .simulate import Simulator, forward
from .transform import CompositeTransform, LinearTransform, Transform, rotation_matrix
