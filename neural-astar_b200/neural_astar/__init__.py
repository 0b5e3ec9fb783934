"""neural_astar — B200-native drop-in for omron-sinicx/neural-astar's planner package.

Same import paths as the reference (`neural_astar.planner.{NeuralAstar,VanillaAstar}`,
`neural_astar.planner.differentiable_astar.AstarOutput`, `neural_astar.utils.*`); the inside of
`DifferentiableAstar.forward` runs in hand-written sm_100a kernels behind a C ABI
(include/nastar_b200.h, libnastar_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
