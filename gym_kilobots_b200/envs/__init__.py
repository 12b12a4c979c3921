from .kilobots_env import KilobotsEnv  # noqa: F401
from .yaml_kilobots_env import YamlKilobotsEnv, EnvConfiguration  # noqa: F401
from .direct_control_kilobots_env import DirectControlKilobotsEnv  # noqa: F401
from .kilobots_test_envs import QuadAssemblyKilobotsEnv, QuadPushingEnv, TriangleTestEnv  # noqa: F401
from .vec_env import KilobotsVecEnv  # noqa: F401
