"""GPU-resident state for PyQMD's hot path: ensembles of nuclei, one big nucleon cloud, and
decay-only populations.  PyTorch tensors own the device memory and the streams; all compute
goes through the C ABI of libpyqmd_b200.so (include/pyqmd_b200.h).

The reference (OtsoBear/PyQMD) keeps one nucleus in Python objects and re-uploads it every
sub-step (nuclear_forces.py:190-234).  Here the state never leaves HBM between sub-steps:

  NucleusEnsemble   many independent nuclei, CSR-packed, one (or several small) per thread block;
                    per sub-step: should_decay -> handle_decay slice -> force + integrate
                    (nuclear_sim.py:165-173), n sub-steps fused per launch
  NucleonCloud      one all-pairs system of N nucleons (nuclear_forces.py:236-323 at scale),
                    i-block sharded over ranks with a per-step position all-gather
  DecayPopulation   particle-less nuclei (decay_chains.py:390-421), Monte Carlo of should_decay
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os

import numpy as np
import torch

from . import _lib, nuclides

DEFAULT_STRENGTHS = (150.0, 30.0, 35.0)          # nuclear_forces.py:13-15
DEFAULT_DT = 1.0 / 240.0                         # nuclear_sim.py:59

# readme.md:43-51 (BASELINE config 3) and nuclear_sim.py:494-504
README_ISOTOPES = ((1, 0), (2, 2), (6, 6), (6, 8), (26, 30), (47, 60), (79, 118), (82, 126),
                   (92, 146))
CODE_ISOTOPES = ((1, 2), (2, 3), (6, 8), (8, 9), (26, 33), (47, 61), (79, 119), (82, 127),
                 (92, 146))

_TEMPLATES = None
_TABLE_CACHE = {}


def resolve_device(device) -> torch.device:
    """``torch.device`` with an explicit index ("cuda" means the device that is current NOW)."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _on_device(fn):
    """Run a method with ``self.device`` as the current CUDA device: the C ABI launches on the
    current device's stream, so pointers of another device must never meet it."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        idx = self.device.index
        if idx is None or idx == torch.cuda.current_device():      # the common case: nothing to switch
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


def layout_templates():
    """Reference-generated initial layouts (Nucleus.initialize_particles, particles.py:62-124),
    64 per isotope, FP32, origin (0, 0); see tests/golden/gen_golden.py."""
    global _TEMPLATES
    if _TEMPLATES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data",
                            "layout_templates.npz")
        _TEMPLATES = dict(np.load(path))
    return _TEMPLATES


def device_table(dt_decay: float, device) -> torch.Tensor:
    key = (float(dt_decay), str(device))
    if key not in _TABLE_CACHE:
        while len(_TABLE_CACHE) >= 8:           # dt varies per frame in the app: bounded cache
            _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
        tab = nuclides.build_device_table(dt_decay)
        _TABLE_CACHE[key] = torch.from_numpy(tab.view(np.uint8).copy()).to(device)
    return _TABLE_CACHE[key]


def check_table_range(zn: torch.Tensor):
    """The dense device table covers Z < 128, N < 192 (PYQMD_TABLE_ZDIM / NDIM, every real nuclide);
    the reference's heuristics accept any integers, so refuse what the table cannot follow."""
    if zn.numel() == 0:
        return
    z, n = zn >> 16, zn & 0xFFFF
    if int(z.max()) >= _lib.TABLE_ZDIM or int(n.max()) >= _lib.TABLE_NDIM or int(z.min()) < 0:
        raise ValueError(f"(Z, N) outside the nuclide table: Z < {_lib.TABLE_ZDIM}, N < {_lib.TABLE_NDIM} "
                         f"required, got Z up to {int(z.max())}, N up to {int(n.max())}")


def initial_half_lives(zn: np.ndarray, dt_decay: float, rng: np.random.Generator):
    """Per-nucleus half-life and per-sub-step decay probability at creation
    (nuclear_sim.py:116 -> get_half_life; particles.py:134-144)."""
    T = np.empty(len(zn), np.float64)
    p = np.empty(len(zn), np.float64)
    for v in np.unique(zn):
        z, n = nuclides.zn_unpack(v)
        sel = zn == v
        kind, value, a, b, unit = nuclides.half_life_class(z, n)
        if kind == _lib.HL_BAND:
            u = rng.random(int(sel.sum()))
            Ts = np.array([nuclides.half_life_from_draw(a, b, unit, float(x)) for x in u])
            T[sel] = Ts
            p[sel] = [nuclides.decay_probability(t, dt_decay) for t in Ts]
        else:
            T[sel] = value
            p[sel] = nuclides.decay_probability(value, dt_decay)
    return T, p


from .sharding import (allgather_positions, cloud_chunk, reduce_scatter_forces,  # noqa: E402,F401
                       shard_range)


# =================================================================================================
class NucleusEnsemble:
    """Independent nuclei on one GPU.  Global nucleus ids ``id_base + k`` key the RNG, so results
    do not depend on how an ensemble is split over GPUs."""

    def __init__(self, zn, offsets, counts, pos, vel, is_proton, *, device="cuda",
                 dt_phys=DEFAULT_DT, dt_decay=DEFAULT_DT, strengths=DEFAULT_STRENGTHS, seed=0,
                 id_base=0, half_life=None, p_decay=None, origin=None, decay=True,
                 event_capacity=1 << 20, init_seed=0, keep_force=False):
        _lib.require_cuda()
        self.device = resolve_device(device)
        dev = self.device
        as_t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray)
                                             else a, dtype=dt).contiguous().to(dev)
        self.zn = as_t(zn, torch.int32)
        if decay:
            check_table_range(self.zn)
        self.offsets = as_t(offsets, torch.int64)
        self.count = as_t(counts, torch.int32)
        self.pos = as_t(pos, torch.float32).reshape(-1, 2).contiguous()
        self.vel = as_t(vel, torch.float32).reshape(-1, 2).contiguous()
        self.is_proton = as_t(is_proton, torch.uint8)
        self.force = torch.zeros_like(self.pos) if keep_force else None
        self.n_nuclei = int(self.zn.numel())
        self.dt_phys, self.dt_decay = float(dt_phys), float(dt_decay)
        self.strengths = tuple(float(s) for s in strengths)
        self.seed, self.id_base, self.decay = int(seed), int(id_base), bool(decay)
        self.step_index = 0
        if half_life is None or p_decay is None:
            T, p = initial_half_lives(self.zn.cpu().numpy(), self.dt_decay,
                                      np.random.default_rng(init_seed))
            half_life = T if half_life is None else half_life
            p_decay = p if p_decay is None else p_decay
        self.half_life = as_t(half_life, torch.float64)
        self.p_decay = as_t(p_decay, torch.float64)
        self.origin = None if origin is None else as_t(origin, torch.float64).reshape(-1, 2)
        self.table = device_table(self.dt_decay, dev)
        self.event_capacity = int(event_capacity)
        self.events_buf = torch.zeros(self.event_capacity * _lib.EVENT_DTYPE.itemsize,
                                      dtype=torch.uint8, device=dev)
        self.event_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.mode_counts = torch.zeros(8, dtype=torch.int64, device=dev)
        # size bins: nuclei with the same initial nucleon count share a launch
        caps = self.count.cpu().numpy()
        self.bins = []
        uniq = np.unique(caps)
        for cap in uniq:
            if cap <= 0:
                continue
            idx = np.nonzero(caps == cap)[0].astype(np.int32)
            lst = None if len(uniq) == 1 else torch.from_numpy(idx).to(dev)
            self.bins.append((int(cap), lst, len(idx)))

    # ---------------------------------------------------------------------------------------------
    @classmethod
    def from_templates(cls, isotopes, n_nuclei, *, device="cuda", id_base=0, rotate=True, **kw):
        """Nucleus ``g = id_base + k`` gets isotope ``isotopes[g % len]``, layout template
        ``(g // len) % 64`` and, if ``rotate``, a rigid rotation by 2*pi*((g // (64*len)) % 1024)/1024
        (SURVEY.md section 8d synthetic inputs).  Built on the device."""
        _lib.require_cuda()
        dev = torch.device(device)
        tm = layout_templates()
        m = len(isotopes)
        g = torch.arange(id_base, id_base + n_nuclei, device=dev, dtype=torch.int64)
        iso = g % m
        a_of = torch.tensor([z + n for z, n in isotopes], device=dev, dtype=torch.int64)
        counts = a_of[iso]
        offsets = torch.cumsum(counts, 0) - counts
        total = int(counts.sum().item())
        pos = torch.empty(total, 2, device=dev, dtype=torch.float32)
        isp = torch.empty(total, device=dev, dtype=torch.uint8)
        zn = torch.tensor([nuclides.zn_pack(z, n) for z, n in isotopes], device=dev,
                          dtype=torch.int32)[iso]
        for k, (z, n) in enumerate(isotopes):
            sel = torch.nonzero(iso == k).flatten()
            if sel.numel() == 0:
                continue
            a = z + n
            txy = torch.from_numpy(tm[f"z{z}_n{n}_xy"]).to(dev)          # [64, a, 2]
            tis = torch.from_numpy(tm[f"z{z}_n{n}_isp"]).to(dev)         # [64, a]
            gk = g[sel]
            t_idx = (gk // m) % txy.shape[0]
            chunk = 1 << 18
            for c0 in range(0, sel.numel(), chunk):
                s = slice(c0, c0 + chunk)
                xy = txy[t_idx[s]]                                       # [c, a, 2]
                if rotate:
                    ang = ((gk[s] // (m * txy.shape[0])) % 1024).to(torch.float64) * (
                        2 * math.pi / 1024)
                    co, si = torch.cos(ang).float()[:, None], torch.sin(ang).float()[:, None]
                    xy = torch.stack((xy[..., 0] * co - xy[..., 1] * si,
                                      xy[..., 0] * si + xy[..., 1] * co), -1)
                flat = (offsets[sel[s]][:, None] + torch.arange(a, device=dev)).reshape(-1)
                pos[flat] = xy.reshape(-1, 2)
                isp[flat] = tis[t_idx[s]].reshape(-1)
        vel = torch.zeros_like(pos)
        return cls(zn, offsets, counts.to(torch.int32), pos, vel, isp, device=device,
                   id_base=id_base, **kw)

    @classmethod
    def from_device_layout(cls, isotopes, n_nuclei, *, device="cuda", id_base=0, layout_seed=0,
                           layout_uniforms=None, **kw):
        """Like ``from_templates`` but every nucleus gets its OWN layout, generated on the device by
        the reference's placement algorithm (Nucleus.initialize_particles, particles.py:62-124;
        pyqmd_ensemble_init_layout) instead of one of 64 reference-generated templates.
        ``layout_uniforms``: optional {A: float64 [n_with_A, A, 21]} draws for parity tests."""
        _lib.require_cuda()
        dev = torch.device(device)
        m = len(isotopes)
        g = torch.arange(id_base, id_base + n_nuclei, device=dev, dtype=torch.int64)
        iso = g % m
        a_of = torch.tensor([z + n for z, n in isotopes], device=dev, dtype=torch.int64)
        counts = a_of[iso]
        offsets = torch.cumsum(counts, 0) - counts
        total = int(counts.sum().item())
        zn = torch.tensor([nuclides.zn_pack(z, n) for z, n in isotopes], device=dev,
                          dtype=torch.int32)[iso]
        pos = torch.zeros(total, 2, device=dev, dtype=torch.float32)
        ens = cls(zn, offsets, counts.to(torch.int32), pos, torch.zeros_like(pos),
                  torch.zeros(total, device=dev, dtype=torch.uint8), device=device, id_base=id_base,
                  **kw)
        ens.init_layout(layout_seed, layout_uniforms)
        return ens

    @_on_device
    def init_layout(self, seed=0, uniforms=None):
        """(Re)generate the initial layout of every nucleus on the device (particles.py:62-124)."""
        lib = _lib.lib()
        keep = []
        for cap, lst, n_list in self.bins:
            start = 1.2 * (cap ** (1 / 3)) * 0.7                  # particles.py:64-66
            radii = (C.c_double * 7)(*[start * (i + 1) / 7 for i in range(7)])      # :68
            u = None
            if uniforms is not None:
                u = torch.as_tensor(uniforms[cap], dtype=torch.float64).contiguous().to(self.device)
                assert tuple(u.shape) == (n_list, cap, 21), (tuple(u.shape), (n_list, cap, 21))
                keep.append(u)
            d = self._desc(cap, lst, n_list, None)
            _lib.check(lib.pyqmd_ensemble_init_layout(C.byref(d), radii, _lib.ptr(u), int(seed),
                                                      _lib.current_stream()),
                       "pyqmd_ensemble_init_layout")
        torch.cuda.synchronize(self.device)

    # ---------------------------------------------------------------------------------------------
    def _desc(self, cap, lst, n_list, uniforms):
        d = _lib.EnsembleDesc()
        d.pos, d.vel, d.is_proton = self.pos.data_ptr(), self.vel.data_ptr(), self.is_proton.data_ptr()
        d.force = _lib.ptr(self.force)
        d.offset, d.count = self.offsets.data_ptr(), self.count.data_ptr()
        d.zn, d.half_life, d.p_decay = (self.zn.data_ptr(), self.half_life.data_ptr(),
                                        self.p_decay.data_ptr())
        d.origin = _lib.ptr(self.origin)
        d.centre = None
        d.n_nuclei, d.id_base = self.n_nuclei, self.id_base
        d.list, d.n_list = _lib.ptr(lst), n_list
        d.cap, d.decay_enabled = cap, 1 if self.decay else 0
        d.strong, d.coulomb, d.pauli = self.strengths
        d.dt_phys, d.dt_decay = self.dt_phys, self.dt_decay
        d.table = self.table.data_ptr()
        d.uniforms, d.uniforms_n = _lib.ptr(uniforms), self.n_nuclei
        d.seed, d.step0 = self.seed, self.step_index
        d.events, d.event_capacity = self.events_buf.data_ptr(), self.event_capacity
        d.event_count, d.mode_counts = self.event_count.data_ptr(), self.mode_counts.data_ptr()
        return d

    @_on_device
    def set_dt_decay(self, dt_decay):
        """Change the dt ``should_decay`` sees (nuclear_sim.py:165: it varies from frame to frame
        with the time scale).  Per-nucleus probabilities are recomputed on the host from the
        current half-lives with the reference's expression (particles.py:134-144) and libm, one
        evaluation per distinct half-life, so decisions stay bit-exact."""
        dt_decay = float(dt_decay)
        if dt_decay == self.dt_decay:
            return
        self.dt_decay = dt_decay
        self.table = device_table(dt_decay, self.device)
        p = nuclides.decay_probabilities(self.half_life.cpu().numpy(), dt_decay)
        self.p_decay.copy_(torch.from_numpy(p))

    @_on_device
    def launch(self, cap, lst_ptr, n_list, n_steps, stream, uniforms=None):
        d = self._desc(cap, None, n_list, uniforms)
        d.list = lst_ptr
        _lib.check(_lib.lib().pyqmd_ensemble_step(C.byref(d), n_steps, stream), "pyqmd_ensemble_step")

    @_on_device
    def step(self, n_steps=1, uniforms=None):
        """``n_steps`` sub-steps of every nucleus (nuclear_sim.py:161-173).  ``uniforms``:
        optional float64 tensor [n_steps, n_nuclei, 4] of draws (slots of SURVEY.md section 8a)
        replacing the Philox stream -- the bit-exact parity path."""
        if uniforms is not None:
            uniforms = torch.as_tensor(uniforms, dtype=torch.float64).contiguous().to(self.device)
            assert tuple(uniforms.shape) == (n_steps, self.n_nuclei, 4)
        lib = _lib.lib()
        stream = _lib.current_stream()
        for cap, lst, n_list in self.bins:
            d = self._desc(cap, lst, n_list, uniforms)
            _lib.check(lib.pyqmd_ensemble_step(C.byref(d), n_steps, stream), "pyqmd_ensemble_step")
        self.step_index += n_steps
        return len(self.bins)

    @_on_device
    def resolve_overlaps(self, uniforms=None):
        """Per-frame projection NuclearSimulation.resolve_overlaps (nuclear_sim.py:355-379) for
        every nucleus.  ``uniforms``: optional float64 [n_nuclei, k] draws for the degenerate
        coincident-pair case.  Returns the device counter tensor of pushes applied so far."""
        if not hasattr(self, "push_count"):
            self.push_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        k = 0
        if uniforms is not None:
            uniforms = torch.as_tensor(uniforms, dtype=torch.float64).contiguous().to(self.device)
            assert uniforms.shape[0] == self.n_nuclei
            k = int(uniforms.shape[1])
        lib = _lib.lib()
        for cap, lst, n_list in self.bins:
            d = self._desc(cap, lst, n_list, None)
            _lib.check(lib.pyqmd_resolve_overlaps(C.byref(d), _lib.ptr(uniforms), k,
                                                  self.push_count.data_ptr(),
                                                  _lib.current_stream()), "pyqmd_resolve_overlaps")
        return self.push_count

    def frame(self, n_substeps=4, uniforms=None):
        """One app frame (nuclear_sim.py:161-176): ``n_substeps`` sub-steps, then the overlap
        projection."""
        self.step(n_substeps, uniforms)
        self.resolve_overlaps()

    @_on_device
    def census(self, sample=None):
        """Branch census of the pair law (device kernel) over ``sample`` (iterable of nucleus
        indices, default all).  Returns (counts dict, algorithmic FLOPs per ordered pair by the
        convention of SURVEY.md section 8d)."""
        counts = torch.zeros(8, dtype=torch.int64, device=self.device)
        lib = _lib.lib()
        if sample is None:
            todo = [(cap, lst, n) for cap, lst, n in self.bins]
            keep = []
        else:
            lst = torch.as_tensor(list(sample), dtype=torch.int32, device=self.device)
            todo, keep = [(1024, lst, int(lst.numel()))], [lst]
        for cap, lst, n_list in todo:
            d = self._desc(cap, lst, n_list, None)
            _lib.check(lib.pyqmd_ensemble_census(C.byref(d), counts.data_ptr(),
                                                 _lib.current_stream()), "pyqmd_ensemble_census")
        c = counts.cpu().tolist()
        names = ("evaluated", "skipped", "hard", "core", "attr", "tail", "pp", "pauli")
        out = dict(zip(names, c))
        flops = (15 * c[0] + 5 * c[2] + 4 * c[3] + 7 * c[4] + 8 * c[5] + 3 * c[6] + 6 * c[7])
        pairs = c[0] + c[1]
        return out, (flops / pairs if pairs else 0.0)

    def pairs_per_step(self):
        """Ordered pair interactions one sub-step evaluates: sum of A(A-1)."""
        c = self.count.to(torch.int64)
        return int((c * (c - 1)).sum().item())

    def events(self):
        """Decay events so far as a structured numpy array sorted by (step, nucleus)."""
        n = min(int(self.event_count.item()), self.event_capacity)
        raw = self.events_buf[: n * _lib.EVENT_DTYPE.itemsize].cpu().numpy()
        ev = raw.view(_lib.EVENT_DTYPE).copy()
        return ev[np.lexsort((ev["nucleus"], ev["step"]))]

    def to_host(self):
        return dict(pos=self.pos.cpu().numpy(), vel=self.vel.cpu().numpy(),
                    is_proton=self.is_proton.cpu().numpy(), offsets=self.offsets.cpu().numpy(),
                    count=self.count.cpu().numpy(), zn=self.zn.cpu().numpy(),
                    half_life=self.half_life.cpu().numpy(), p_decay=self.p_decay.cpu().numpy(),
                    mode_counts=self.mode_counts.cpu().numpy())


# =================================================================================================
class HostEnsembleRunner:
    """End-to-end path with HOST-resident state, the shape of the reference's per-step call
    (nuclear_forces.py:190-234: pack -> H2D -> kernel -> D2H -> write back): the ensemble lives in
    pinned host memory; every ``step`` uploads it, runs the sub-steps and downloads it again -- one
    C-ABI call, pyqmd_ensemble_step_host, which cuts the nuclei into chunks and streams them through
    three in-order lanes (upload, compute, download), so both PCIe directions and the kernels overlap."""

    def __init__(self, ens: NucleusEnsemble, chunks=8):
        self.ens = ens
        self.device = ens.device
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        self.h_pos, self.h_vel, self.h_isp = pin(ens.pos), pin(ens.vel), pin(ens.is_proton)
        self.h_count, self.h_zn = pin(ens.count), pin(ens.zn)
        off = ens.offsets.cpu().numpy()
        cnt = ens.count.cpu().numpy()
        n = ens.n_nuclei
        chunks = max(1, min(chunks, n))
        bounds = [round(k * n / chunks) for k in range(chunks + 1)]
        dev = ens.device
        assert len(ens.bins) <= _lib.MAX_CHUNK_LAUNCHES
        self.lists = []          # per size bin: (cap, device index list, host copy)
        for cap, lst, n_list in ens.bins:
            idx = (np.arange(n, dtype=np.int32) if lst is None else lst.cpu().numpy())
            self.lists.append((cap, torch.from_numpy(idx).to(dev), idx))
        rows = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b <= a:
                continue
            c = _lib.HostChunk()
            c.nuc0, c.nuc1 = a, b
            c.slot0, c.slot1 = int(off[a]), int(off[b - 1] + cnt[b - 1])
            k = 0
            for cap, lst_dev, idx in self.lists:
                lo, hi = np.searchsorted(idx, a), np.searchsorted(idx, b)
                if hi > lo:
                    c.cap[k], c.list[k], c.n_list[k] = cap, lst_dev.data_ptr() + 4 * int(lo), int(hi - lo)
                    k += 1
            c.n_launch = k
            rows.append(c)
        self.n_chunks = len(rows)
        self.chunk_array = (_lib.HostChunk * self.n_chunks)(*rows)
        # without decay the nucleon types never change: they are uploaded with the first step only
        self.types_resident = False
        self.h2d_bytes = int(self.h_pos.nbytes + self.h_vel.nbytes) + (int(self.h_isp.nbytes) if ens.decay else 0)
        # without decay the types / counts / (Z, N) cannot change: only positions and velocities
        # come back, like the reference's download (nuclear_forces.py:227)
        self.d2h_bytes = int(self.h_pos.nbytes + self.h_vel.nbytes) + (
            int(self.h_isp.nbytes + self.h_count.nbytes + self.h_zn.nbytes) if ens.decay else 0)

    @_on_device
    def step(self, n_steps=1):
        ens = self.ens
        d = ens._desc(1, None, 0, None)
        isp_ptr = None if (self.types_resident and not ens.decay) else self.h_isp.data_ptr()
        self.types_resident = True
        _lib.check(_lib.lib().pyqmd_ensemble_step_host(
            C.byref(d), self.h_pos.data_ptr(), self.h_vel.data_ptr(), isp_ptr,
            self.h_count.data_ptr(), self.h_zn.data_ptr(), self.chunk_array, self.n_chunks, n_steps,
            _lib.current_stream()), "pyqmd_ensemble_step_host")
        ens.step_index += n_steps


# =================================================================================================
class NucleonCloud:
    """One N-nucleon system on ``world`` GPUs.  Every rank holds a full replica of the positions and
    owns the i-block [i0, i1) for the integration; new positions are all-gathered after each step.

    scheme="symmetric" (default): every unordered pair is evaluated once (the law is symmetric);
        i-block rows are dealt to the ranks, each rank adds its share of the forces on ALL nucleons
        into int64 fixed-point accumulators, an integer reduce-scatter delivers every rank the
        exact, order-independent totals of its own block.
    scheme="ordered": rank r evaluates the ordered pairs (i, j) for its own i only (the
        decomposition BASELINE.json names: i-block + position all-gather, no force exchange).

    exchange (symmetric scheme, world > 1):
      "peer" (default): accumulators and position replicas live in symmetric memory
        (torch.distributed._symmetric_memory: every rank maps every peer's buffers over NVLink);
        ONE kernel per step pulls + sums + clears the accumulators of the rank's block from all
        peers, integrates, and pushes the new positions into all replicas
        (pyqmd_cloud_exchange_integrate), bracketed by two device-side barriers.  No NCCL call on
        the data path.  Raises if symmetric memory cannot be set up, unless
        ``allow_nccl_fallback=True`` (then "nccl" is used, with a warning); ``self.exchange`` names
        the exchange that is really in use.
      "nccl": reduce_scatter_tensor(int64 SUM) -> integrate -> all_gather_into_tensor.
    Both exchanges give bit-identical results (integer sums).
    """

    def __init__(self, pos, is_proton, vel=None, *, device="cuda", dt=DEFAULT_DT,
                 strengths=DEFAULT_STRENGTHS, rank=0, world=1, group=None, sort=True,
                 keep_force=False, scheme="symmetric", exchange="peer", allow_nccl_fallback=False,
                 skip_exact_zeros=False):
        _lib.require_cuda()
        assert scheme in ("symmetric", "ordered") and exchange in ("peer", "nccl")
        dev = self.device = resolve_device(device)
        pos = torch.as_tensor(pos, dtype=torch.float32).reshape(-1, 2).to(dev)
        isp = torch.as_tensor(is_proton, dtype=torch.uint8).to(dev)
        vel = torch.zeros_like(pos) if vel is None else torch.as_tensor(
            vel, dtype=torch.float32).reshape(-1, 2).to(dev)
        self.n = int(pos.shape[0])
        self.dt, self.strengths = float(dt), tuple(float(s) for s in strengths)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.scheme = scheme
        # opt-in: skip tiles / exponentials whose contribution is exactly zero in FP32 (bit-identical
        # results, see PYQMD_CLOUD_SKIP_EXACT_ZEROS in include/pyqmd_b200.h)
        self.pair_flags = _lib.CLOUD_SKIP_EXACT_ZEROS if skip_exact_zeros else 0
        self.chunk = cloud_chunk(self.n, self.world)
        self.i0, self.i1 = shard_range(self.n, self.rank, self.world)
        self.perm = None
        if sort and self.n > 0:
            self.perm = self._sort_perm(pos, isp)
            pos, vel, isp = pos[self.perm].contiguous(), vel[self.perm].contiguous(), isp[self.perm].contiguous()
        padded = self.chunk * self.world
        self.exchange = exchange if (self.world > 1 and scheme == "symmetric") else None
        self._symm = None
        if self.exchange == "peer":
            try:
                self._setup_peer_buffers(padded)
            except Exception as exc:             # noqa: BLE001
                if not allow_nccl_fallback:      # opt-in only: a silent fallback would leave the fused
                    raise RuntimeError(          # peer-memory kernel untested without anyone noticing
                        f"symmetric memory unavailable ({exc!r}); pass exchange='nccl' or "
                        "allow_nccl_fallback=True to use the NCCL exchange") from exc
                import warnings
                warnings.warn(f"symmetric memory unavailable ({exc!r}); using the NCCL exchange")
                self.exchange, self._symm = "nccl", None
        if self._symm is None:
            self.pos = torch.zeros(padded, 2, device=dev, dtype=torch.float32)
            self.pos_next = torch.zeros(padded, 2, device=dev, dtype=torch.float32)
        self.pos[: self.n] = pos
        self.vel = vel.contiguous()
        self.is_proton = isp.contiguous()
        self.force = torch.zeros(self.n, 2, device=dev, dtype=torch.float32) if keep_force else None
        ws = int(_lib.lib().pyqmd_cloud_workspace_bytes(self.n))
        self.workspace = torch.zeros(max(ws, 64), dtype=torch.uint8, device=dev)
        # fixed-point force accumulators of the symmetric scheme (kept zero between steps)
        if self._symm is None:
            self.acc = self.acc_mine = None
            if scheme == "symmetric":
                self.acc = torch.zeros(padded, 2, device=dev, dtype=torch.int64)
                self.acc_mine = (torch.zeros(self.chunk, 2, device=dev, dtype=torch.int64)
                                 if self.world > 1 else self.acc)
        self.force_scale_log2 = int(_lib.lib().pyqmd_cloud_force_scale_log2(max(self.n, 1)))
        self.steps_done = 0
        self.profile = None              # set to [] to collect (start, end) events of the pair-force launches

    def pair_kernel_ms(self):
        """Mean device time of this rank's pair-force launches collected in ``self.profile``."""
        torch.cuda.synchronize(self.device)
        t = [a.elapsed_time(b) for a, b in self.profile]
        return sum(t) / max(len(t), 1)

    def _setup_peer_buffers(self, padded):
        """acc + both position replicas in symmetric memory; device arrays of the peers' pointers."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        dev = self.device
        bufs, hdls = {}, {}
        for name, dtype in (("acc", torch.int64), ("pos_a", torch.float32), ("pos_b", torch.float32)):
            t = symm_mem.empty(padded, 2, dtype=dtype, device=dev)
            hdl = symm_mem.rendezvous(t, group)
            t.zero_()
            bufs[name], hdls[name] = t, hdl
        assert hdls["acc"].world_size == self.world and hdls["acc"].rank == self.rank
        self.acc, self.acc_mine = bufs["acc"], None
        self.pos, self.pos_next = bufs["pos_a"], bufs["pos_b"]
        ptrs = lambda h: torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=dev)
        self._symm = {"hdl": hdls, "acc_ptrs": ptrs(hdls["acc"]),
                      "pos_ptrs": {bufs["pos_a"].data_ptr(): ptrs(hdls["pos_a"]),
                                   bufs["pos_b"].data_ptr(): ptrs(hdls["pos_b"])}}
        torch.cuda.synchronize(dev)
        hdls["acc"].barrier(channel=0)           # every rank's buffers are zeroed before the first step

    @_on_device
    def _sort_perm(self, pos, isp):
        lo = pos.min(0).values
        extent = float((pos.max(0).values - lo).max().item()) * 1.0001 + 1e-6
        keys = torch.empty(self.n, dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().pyqmd_cloud_sort_keys(
            pos.contiguous().data_ptr(), isp.data_ptr(), self.n, float(lo[0]), float(lo[1]),
            extent, keys.data_ptr(), _lib.current_stream()), "pyqmd_cloud_sort_keys")
        # stable: every rank sorts its own replica, equal keys must come out in the same order
        return torch.argsort(keys, stable=True)

    @_on_device
    def step(self, n_steps=1):
        lib = _lib.lib()
        S, Cc, P = self.strengths
        for _ in range(n_steps):
            stream = _lib.current_stream()
            if self.acc is None:
                # ordered i-block scheme: forces on [i0, i1) from all j, no force exchange
                _lib.check(lib.pyqmd_cloud_step(
                    self.pos.data_ptr(), self.pos_next.data_ptr(), self.vel.data_ptr(),
                    _lib.ptr(self.force), self.is_proton.data_ptr(), self.n, self.i0, self.i1, S, Cc,
                    P, self.dt, self.workspace.data_ptr(), stream), "pyqmd_cloud_step")
            else:
                if self.profile is not None:     # per-rank kernel time (load-balance evidence)
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                _lib.check(lib.pyqmd_cloud_pair_forces_ex(
                    self.pos.data_ptr(), self.is_proton.data_ptr(), self.n, self.rank, self.world,
                    S, Cc, P, self.acc.data_ptr(), self.workspace.data_ptr(), self.pair_flags, stream),
                    "pyqmd_cloud_pair_forces")
                if self.profile is not None:
                    ev[1].record()
                    self.profile.append(ev)
                if self._symm is not None:
                    # peer-memory exchange fused with the integration (no NCCL on the data path)
                    hdl = self._symm["hdl"]["acc"]
                    hdl.barrier(channel=0)       # all pair-force kernels finished
                    _lib.check(lib.pyqmd_cloud_exchange_integrate(
                        self.pos.data_ptr(), self.vel.data_ptr(), _lib.ptr(self.force), self.n,
                        self.i0, self.i1, self.dt, self._symm["acc_ptrs"].data_ptr(),
                        self._symm["pos_ptrs"][self.pos_next.data_ptr()].data_ptr(), self.world,
                        self.workspace.data_ptr(), _lib.current_stream()),
                        "pyqmd_cloud_exchange_integrate")
                    hdl.barrier(channel=1)       # all pushes landed, all consumed entries cleared
                    self.pos, self.pos_next = self.pos_next, self.pos
                    self.steps_done += 1
                    continue
                if self.world > 1:
                    reduce_scatter_forces(self.acc, self.acc_mine, self.rank, self.world, self.group)
                    self.acc.zero_()
                if self.i1 > self.i0:
                    _lib.check(lib.pyqmd_cloud_integrate(
                        self.pos.data_ptr(), self.pos_next.data_ptr(), self.vel.data_ptr(),
                        _lib.ptr(self.force), self.n, self.i0, self.i1, self.dt,
                        self.acc_mine.data_ptr() + (16 * self.i0 if self.world == 1 else 0),
                        self.workspace.data_ptr(), _lib.current_stream()),
                        "pyqmd_cloud_integrate")
            allgather_positions(self.pos_next, self.rank, self.world, self.chunk, self.group)
            self.pos, self.pos_next = self.pos_next, self.pos
            self.steps_done += 1

    # -- host-resident state (multi-GPU counterpart of pyqmd_cloud_step_host) ---------------------
    @_on_device
    def download_block(self, h_pos, h_vel):
        """This rank's block [i0, i1) (sorted order) -> pinned host tensors [i1 - i0, 2]."""
        h_pos.copy_(self.pos[self.i0:self.i1], non_blocking=True)
        h_vel.copy_(self.vel[self.i0:self.i1], non_blocking=True)
        torch.cuda.synchronize(self.device)

    @_on_device
    def step_host(self, h_pos, h_vel):
        """One step with the state of this rank's block in (pinned) HOST memory, the shape of the
        reference's per-step call (nuclear_forces.py:190-234) on several GPUs: upload the block,
        all-gather the positions into every replica, step, download the block.  Blocking."""
        self.pos[self.i0:self.i1].copy_(h_pos, non_blocking=True)
        self.vel[self.i0:self.i1].copy_(h_vel, non_blocking=True)
        allgather_positions(self.pos, self.rank, self.world, self.chunk, self.group)
        self.step(1)
        self.download_block(h_pos, h_vel)

    def pairs_per_step(self):
        """Ordered pairs this rank evaluates per step: (i1 - i0) * (n - 1)."""
        return (self.i1 - self.i0) * (self.n - 1)

    def _unsort(self, t):
        if self.perm is None:
            return t
        out = torch.empty_like(t)
        out[self.perm] = t
        return out

    def positions(self):
        """Positions in the caller's original nucleon order."""
        return self._unsort(self.pos[: self.n])

    def velocities(self):
        return self._unsort(self.vel)

    def forces(self):
        return None if self.force is None else self._unsort(self.force)


# =================================================================================================
class DecayPopulation:
    """Particle-less nuclei: per sub-step should_decay (decay_chains.py:400-421) and, on a hit,
    the (Z, N) / half-life update of handle_decay (nuclear_sim.py:213,288-289,353).

    ``half_life`` / ``p_decay``: optional caller-chosen per-nucleus values (then the kernel reads them
    for every nucleus, 20 B per nucleus and launch); by default they follow get_half_life and the
    kernel takes the values of tabulated nuclides from the nuclide table (4 B per nucleus and launch)."""

    COUNT_POOL_ROWS = 4096

    def __init__(self, zn, *, device="cuda", dt_decay, seed=0, id_base=0, watch=(),
                 half_life=None, p_decay=None, init_seed=0):
        _lib.require_cuda()
        dev = self.device = resolve_device(device)
        self.zn = torch.as_tensor(zn, dtype=torch.int32).contiguous().to(dev)
        check_table_range(self.zn)
        self.n = int(self.zn.numel())
        self.dt_decay, self.seed, self.id_base = float(dt_decay), int(seed), int(id_base)
        self.per_nucleus_state = half_life is not None or p_decay is not None
        if half_life is None or p_decay is None:
            uniq, inv = torch.unique(self.zn, return_inverse=True)
            Tu, pu = initial_half_lives(uniq.cpu().numpy(), self.dt_decay,
                                        np.random.default_rng(init_seed))
            kinds = [nuclides.half_life_class(*nuclides.zn_unpack(v))[0] for v in uniq.tolist()]
            if any(k == _lib.HL_BAND for k in kinds):
                T, p = initial_half_lives(self.zn.cpu().numpy(), self.dt_decay,
                                          np.random.default_rng(init_seed))
                T, p = torch.from_numpy(T), torch.from_numpy(p)
            else:
                T = torch.from_numpy(Tu).to(dev)[inv]
                p = torch.from_numpy(pu).to(dev)[inv]
            half_life = T if half_life is None else half_life
            p_decay = p if p_decay is None else p_decay
        self.half_life = torch.as_tensor(half_life, dtype=torch.float64).contiguous().to(dev)
        self.p_decay = torch.as_tensor(p_decay, dtype=torch.float64).contiguous().to(dev)
        self.table = device_table(self.dt_decay, dev)
        self.watch = [nuclides.zn_pack(z, n) for z, n in watch][:8]
        self.step_index = 0
        self._desc = None
        self._pool, self._pool_used = None, 0

    def _count_rows(self, n_steps):
        """Zeroed [n_steps, 16] view of a pooled counter buffer: one allocation + fill per
        COUNT_POOL_ROWS sub-steps instead of one per call (at 8 GPUs a step is ~30 us of kernel)."""
        if n_steps > self.COUNT_POOL_ROWS:
            return torch.zeros(n_steps, _lib.COUNT_COLS, dtype=torch.int64, device=self.device)
        if self._pool is None or self._pool_used + n_steps > self.COUNT_POOL_ROWS:
            self._pool = torch.zeros(self.COUNT_POOL_ROWS, _lib.COUNT_COLS, dtype=torch.int64,
                                     device=self.device)
            self._pool_used = 0
        out = self._pool[self._pool_used:self._pool_used + n_steps]
        self._pool_used += n_steps
        return out

    def bytes_per_nucleus_launch(self):
        """(HBM bytes the kernel reads per nucleus and launch, explanation) for the roofline."""
        if self.per_nucleus_state:
            return 20, "zn + caller-supplied half-life + p (20 B); written back only for decayed nuclei"
        return 4, ("zn only (4 B): half-life and p of tabulated nuclides come from the cached table row; "
                   "per-nucleus side arrays are read for estimated half-lives only, written for decayed nuclei")

    @_on_device
    def step(self, n_steps=1, uniforms=None, want_decisions=False):
        """Returns (counts[n_steps, 16] int64 tensor, decisions[n_steps, n] uint8 or None);
        counts columns: decays by DecayType value 0..7, then decays of watch[k] in 8..15."""
        dev = self.device
        counts = self._count_rows(n_steps)
        decided = torch.zeros(n_steps, self.n, dtype=torch.uint8, device=dev) if want_decisions else None
        if uniforms is not None:
            uniforms = torch.as_tensor(uniforms, dtype=torch.float64).contiguous().to(dev)
            assert tuple(uniforms.shape) == (n_steps, self.n, 4)
        if self._desc is None:
            d = self._desc = _lib.PopulationDesc()
            d.zn, d.half_life, d.p_decay = self.zn.data_ptr(), self.half_life.data_ptr(), self.p_decay.data_ptr()
            d.n, d.id_base, d.table, d.dt_decay = self.n, self.id_base, self.table.data_ptr(), self.dt_decay
            d.seed, d.n_watch = self.seed, len(self.watch)
            d.flags = _lib.POP_PER_NUCLEUS_STATE if self.per_nucleus_state else 0
            for k, v in enumerate(self.watch):
                d.watch_zn[k] = v
        d = self._desc
        d.uniforms, d.uniforms_n = _lib.ptr(uniforms), self.n
        d.step0 = self.step_index
        d.step_counts, d.decided = counts.data_ptr(), _lib.ptr(decided)
        _lib.check(_lib.lib().pyqmd_population_step(C.byref(d), n_steps, _lib.current_stream()),
                   "pyqmd_population_step")
        self.step_index += n_steps
        return counts, decided
