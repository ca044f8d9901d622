"""`import drone_env` -> the B200-native drop-in (see scalable_collision_avoidance_rl_b200/drone_env.py).

With this repository's root ahead of the reference on sys.path, the reference's
train_problem.py / benchmark_agent.py import this module instead of their own.
"""
from scalable_collision_avoidance_rl_b200.drone_env import *  # noqa: F401,F403
from scalable_collision_avoidance_rl_b200.drone_env import (  # noqa: F401
    dim, dt, max_time_steps, drones, gradient_control, proportional_control, running_average,
    plot_rewards, plot_grads, num_to_rgb)
