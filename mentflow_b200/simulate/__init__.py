from .simulate import Simulator, forward
from .transform import (CompositeTransform, LinearTransform, MultipoleTransform, ProjectionTransform, Transform,
                        reverse_momentum, rotation_matrix)
