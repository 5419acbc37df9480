"""pyqmd_b200 -- B200-native implementation of PyQMD's data-parallel hot path
(all-pairs nucleon force -> damped Euler integrate -> per-nucleus stochastic decay).

Reference-facing interface (same names as OtsoBear/PyQMD):
    NuclearForces                                   nuclear_forces.py:10
    ParticleType, DecayType, Particle, Nucleus      particles.py:5-60
    get_decay_product, get_half_life, HALF_LIVES, DECAY_CHAINS, create_*   decay_chains.py
GPU-resident state (new; the reference never batches):
    NucleusEnsemble, NucleonCloud, DecayPopulation, HostEnsembleRunner   pyqmd_b200/state.py
    HeadlessSimulation (frame driver, nuclear_sim.py:118-176 without pygame)  pyqmd_b200/sim.py
All compute goes through the C ABI in include/pyqmd_b200.h (libpyqmd_b200.so, hand-written
sm_100a CUDA); there is no CPU fallback.
"""
from .types import DecayType, Nucleus, Particle, ParticleType  # noqa: F401
from .nuclides import (DECAY_CHAINS, HALF_LIVES, create_alpha, create_beta_minus,  # noqa: F401
                       create_beta_plus, create_fission, create_gamma, create_neutron,
                       create_proton, expand_decay_chain, get_decay_product, get_half_life)
from .forces import NuclearForces  # noqa: F401


def __getattr__(name):
    # torch-dependent state classes are imported lazily so the reference-shaped API above
    # stays importable in a few milliseconds
    if name in ("NucleusEnsemble", "NucleonCloud", "DecayPopulation", "HostEnsembleRunner",
                "README_ISOTOPES", "CODE_ISOTOPES", "shard_range"):
        from . import state
        return getattr(state, name)
    if name in ("HeadlessSimulation", "TIME_SCALE_PRESETS"):
        from . import sim
        return getattr(sim, name)
    raise AttributeError(name)


__version__ = "0.1.0"
