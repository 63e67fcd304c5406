"""B200-native Soft Actor-Critic update engine behind the Python surface of dogeplusplus/torch-actor-critic.

Module paths mirror the reference so that its call sites read the same:

    from torch_actor_critic_b200.sac.algorithm import SAC, update_targets
    from torch_actor_critic_b200.buffer.replay_buffer import ReplayBuffer, Batch
    from torch_actor_critic_b200.networks.linear import Actor, Critic, DoubleCritic

Everything numerical runs in `libsacb200.so` (hand-written sm_100a CUDA, C ABI in include/sacb200.h, built by
`__graft_entry__.build()`).  There is no CPU or PyTorch fallback: importing `_lib` without the built library, or
creating an engine without a CUDA device, raises.
"""
__version__ = "0.1.0"
