"""Randomised differential check of the ensemble kernel against the CPU oracle: many small random
nuclei (sizes 1..260, clustered / spread / with coincident nucleons), random strengths and dt."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from gpu_util import extent_of, force_error, oracle_step
from pyqmd_b200.state import NucleusEnsemble



def run(seed=0, trials=40):
  rng = np.random.default_rng(seed)
  worst, worst_f, n_cases, n_amb, detail = 0.0, 0.0, 0, 0, None
  for trial in range(trials):
      S, Cq, P = float(rng.uniform(0, 400)), float(rng.uniform(0, 80)), float(rng.uniform(0, 80))
      dt = float(rng.choice([1 / 240, 1 / 60, 1e-3]))
      sizes = rng.integers(1, 261, 48)
      pos, vel, isp, off, cnt, zn = [], [], [], [], [], []
      o = 0
      for a in sizes:
          spread = float(rng.choice([0.5, 2.0, 6.0, 20.0]))
          p = (rng.normal(0, spread, (a, 2))).astype(np.float32)
          if a > 3 and rng.random() < 0.3:
              p[1] = p[0]                      # coincident pair
              p[2] = p[0] + np.float32(0.05)   # inside the d2 < 0.01 skip
          t = (rng.random(a) < 0.4).astype(np.uint8)
          pos.append(p); vel.append(rng.normal(0, 1, (a, 2)).astype(np.float32)); isp.append(t)
          off.append(o); cnt.append(a); zn.append((int(t.sum()) << 16) | int(a - t.sum())); o += a
      ens = NucleusEnsemble(np.array(zn, np.int32), np.array(off, np.int64), np.array(cnt, np.int32),
                            np.concatenate(pos), np.concatenate(vel), np.concatenate(isp), decay=False,
                            dt_phys=dt, strengths=(S, Cq, P), keep_force=True)
      ens.step(1)
      got = ens.pos.cpu().numpy()
      gotf = ens.force.cpu().numpy()
      for k, a in enumerate(sizes):
          ox, oy, _, _, fx, fy, amb = oracle_step(pos[k], vel[k], isp[k], dt, S, Cq, P)
          ok = ~amb
          if ok.any() and np.hypot(fx[ok], fy[ok]).max() > 1e-3:
              worst_f = max(worst_f, force_error(gotf[off[k]:off[k] + a], fx, fy, amb))
          n_amb += int(amb.sum())
          if ok.any():
              ext = max(extent_of(pos[k]), 1.0)      # A = 1 has no extent: absolute FP32 rounding then
              g = got[off[k]:off[k] + a]
              e = np.hypot(g[:, 0] - ox, g[:, 1] - oy) * ok
              err = float(e.max() / ext)
              if err > worst:
                  i = int(e.argmax())
                  d = np.hypot(pos[k][:, 0].astype(np.float64) - float(pos[k][i, 0]),
                               pos[k][:, 1].astype(np.float64) - float(pos[k][i, 1]))
                  d[i] = 1e9
                  detail = dict(trial=trial, S=S, C=Cq, P=P, dt=dt, A=int(a), ext=ext, nucleon=i,
                                abs_err=float(e.max()), nearest=float(d.min()), n_within_0p2=int((d < 0.2).sum()),
                                is_proton=int(isp[k][i]))
              worst = max(worst, err)
          n_cases += 1
  return {"nuclei": n_cases, "worst_pos_err": worst, "worst_force_err_l2": worst_f, "ambiguous_nucleons": n_amb,
          "worst_case": detail}


if __name__ == "__main__":
    res = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    print(json.dumps(res))
    assert res["worst_pos_err"] <= 1e-5, res["worst_pos_err"]
    assert res["worst_force_err_l2"] <= 1e-5, res["worst_force_err_l2"]
