from tracktolearn_b200.environments.tracking_env import TrackingEnvironment  # noqa: F401
from tracktolearn_b200.environments.noisy_tracking_env import NoisyTrackingEnvironment  # noqa: F401

# the name used by the task statement; the reference class is TrackingEnvironment.
RLTrackingEnvironment = TrackingEnvironment
