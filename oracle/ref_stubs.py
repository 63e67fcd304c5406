"""Import stubs that let the UNMODIFIED reference modules import without gym / mlflow / mpi4py / dm_control.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by oracle/make_golden.py (golden vectors from
/root/reference) and by bench.py's `--impl reference` arm (the reference copy under oracle/_ref, see oracle/build_ref.py).

With one rank, `mpi_avg_grads` / `sync_params` early-return (sac/mpi.py:79-80,94-95), so the MPI stub needs nothing
beyond rank 0 / size 1.  `gym.spaces.Box`, `gym.envs.registration.register` and `dm_control...basic_cmu_2019` are only
touched at import time by environments/ (pulled in by networks/convolutional.py:11).
"""
import sys
import types


def install_stubs():
    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not getattr(m, "__sacb_stub__", False):
            m = types.ModuleType(name)
            m.__sacb_stub__ = True
            sys.modules[name] = m
        m.__dict__.update(attrs)
        return m

    class Box:                                           # gym.spaces.Box(low, high, shape): only shape/high are read (main.py:57-59)
        def __init__(self, low, high, shape=None, dtype=None):
            import numpy as np
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=np.float32)
            self.high = np.full(self.shape, high, dtype=np.float32)

        def sample(self):
            import numpy as np
            return np.random.uniform(self.low, self.high).astype(np.float32)

    gym = stub("gym", Env=object)
    gym.spaces = stub("gym.spaces", Box=Box)
    gym.envs = stub("gym.envs")
    gym.envs.registration = stub("gym.envs.registration", register=lambda **kw: None)
    ml = stub("mlflow")
    ml.pytorch = stub("mlflow.pytorch")

    class _Comm:
        def Get_rank(self):
            return 0

        def Get_size(self):
            return 1

    mpi = types.SimpleNamespace(COMM_WORLD=_Comm(), SUM=0, MIN=1, MAX=2)
    stub("mpi4py", MPI=mpi)
    stub("mpi4py.MPI", **mpi.__dict__)
    dm = stub("dm_control")
    dm.locomotion = stub("dm_control.locomotion")
    dm.locomotion.examples = stub("dm_control.locomotion.examples")
    dm.locomotion.examples.basic_cmu_2019 = stub("dm_control.locomotion.examples.basic_cmu_2019")


def import_reference(root):
    """Import the reference's hot-path modules from `root` (/root/reference or oracle/_ref), unmodified."""
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import sac.algorithm as alg
    from buffer.replay_buffer import Batch, ReplayBuffer
    from networks.linear import Actor, Critic, DoubleCritic
    return alg, Batch, ReplayBuffer, Actor, Critic, DoubleCritic


def import_reference_visual(root):
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import sac.algorithm as alg
    from buffer.visual_replay_buffer import VisualBatch, VisualReplayBuffer
    from environments.wall_runner import MultiObservation
    from networks.convolutional import VisualActor, VisualCritic, VisualDoubleCritic, calculate_size, simple_cnn
    return alg, VisualBatch, VisualReplayBuffer, MultiObservation, VisualActor, VisualCritic, VisualDoubleCritic, calculate_size, simple_cnn
