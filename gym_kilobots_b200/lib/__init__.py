from .kilobot import Kilobot, PhototaxisKilobot, SimplePhototaxisKilobot, SimpleVelocityControlKilobot, \
    SimpleAccelerationControlKilobot  # noqa: F401
from .body import Body, Quad, CornerQuad, Triangle, Circle, CForm, TForm, LForm  # noqa: F401
from .light import CircularGradientLight, GradientLight, CompositeLight  # noqa: F401
from .world import World  # noqa: F401
