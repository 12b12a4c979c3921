"""gym_kilobots_b200 -- B200-native batched simulator for the gym-kilobots `KilobotsEnv.step` hot path.

Public surface: `gym_kilobots_b200.envs` (KilobotsVecEnv + the reference's env classes),
`gym_kilobots_b200.lib` (the reference's body / kilobot / light constructors) and
`gym_kilobots_b200.scenarios` (synthetic batches for the BASELINE.json configurations).
"""
__version__ = "0.1.0"
