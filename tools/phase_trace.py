#!/usr/bin/env python
"""Diagnostics (needs a -DS64_TRACE build selected with GCA_LIB_PATH): mean per-env phase timestamps of the
64x64 step kernel, L2-cold (flushed) vs L2-warm.  Not part of the product path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_cellular_automata_b200 import _lib
from gym_cellular_automata_b200.packed import StepOutputs
from gym_cellular_automata_b200._lib import GcaStepOut, ptr
from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv

N = 4096; K = 4
dev = torch.device("cuda", 0)
env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.48, speed_act=0.12, use_hidden=True,
                                     substeps=K, rng_mode="legacy", seed=0, hidden="random", obs_mode="none",
                                     auto_reset=True, collect_stats=True, device=dev, balance_every=8)
env.reset()
o = env._out
o.stats = torch.zeros(8 + 32 * N, dtype=torch.int64, device=dev)
o._c = GcaStepOut(ptr(o.reward).value, ptr(o.step_reward).value, ptr(o.terminated).value, ptr(o.counts).value,
                  ptr(o.obs_night).value, ptr(o.stats).value)
env._version_structs += 1
gen = torch.Generator(device=dev); gen.manual_seed(0)
def act():
    return torch.stack([torch.randint(0, 9, (N,), device=dev, generator=gen), torch.randint(0, 2, (N,), device=dev, generator=gen),
                        torch.randint(0, 3, (N,), device=dev, generator=gen)], -1).to(torch.int32).contiguous()
if os.environ.get('PREROLL'):
    from gym_cellular_automata_b200.workload import stationary_preroll
    stationary_preroll(env, int(os.environ['PREROLL']), 32)
for _ in range(100): env.step_device(act())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["setup+publish key", "barrier+wait+convert+list", "scan", "(rm)"] + sum([[f"s{j} owner-pre", f"s{j} pooled", f"s{j} barrier"] for j in range(K)], [])
for mode in ("cold", "warm", "cold", "warm"):
    a = act()
    if mode == "cold": flush.fill_(1)
    o.stats.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_device(a); e1.record(); torch.cuda.synchronize()
    tr = o.stats[8:].cpu().numpy().reshape(N, 32).astype(np.float64)
    feat = tr[:, 20:24].copy()
    if mode == "cold":
        order = env._state.order.cpu().numpy() if env._state.order is not None else np.arange(N)
        E = int(os.environ.get('E', 14))
        ncta = N // E
        byslot = feat[order[:ncta * E]].reshape(ncta, E, 4)
        T = byslot[:, :, 3].max(1)
        X = np.stack([np.ones(ncta), byslot[:, :, 0].sum(1), byslot[:, :, 1].sum(1), byslot[:, :, 2].sum(1),
                      byslot[:, :, 0].max(1), byslot[:, :, 1].max(1)], 1)
        coef, res, *_ = np.linalg.lstsq(X, T, rcond=None)
        pred = X @ coef
        print("  CTA time: mean %.1fk max %.1fk std %.1fk | fit [1, sum entries, sum pairs, sum rows, max entries, max pairs] = %s, residual std %.1fk"
              % (T.mean() / 1e3, T.max() / 1e3, T.std() / 1e3, np.round(coef, 2).tolist(), (T - pred).std() / 1e3))
        est = byslot[:, :, 0].sum(1) * 2 + byslot[:, :, 1].sum(1)
        print("  corr(T, current estimate sum) = %.3f ; estimate sum mean %.0f std %.0f" % (np.corrcoef(T, est)[0, 1], est.mean(), est.std()))
        full = o.stats[8:].cpu().numpy().reshape(N, 32).astype(np.float64)[order[:ncta * E]].reshape(ncta, E, 32)
        smid = full[:, 0, 24].astype(int); tend = full[:, :, 25].max(1); tend -= tend.min()
        cyc = full[:, :, 23].max(1)
        tstart = tend - cyc / 1.965  # ns
        print("  CTA start (ns rel.): min %.0f p50 %.0f p90 %.0f max %.0f ; end: p50 %.0f max %.0f" % (
            tstart.min(), np.percentile(tstart, 50), np.percentile(tstart, 90), tstart.max(), np.percentile(tend, 50), tend.max()))
        per_sm = {}
        for c in range(ncta): per_sm.setdefault(smid[c], []).append(cyc[c])
        nper = np.array([len(v) for v in per_sm.values()]); msm = np.array([np.mean(v) for v in per_sm.values()])
        ssm = np.array([np.std(v) for v in per_sm.values()])
        print("  SMs used %d, CTAs/SM min %d max %d ; per-SM mean CTA cycles: min %.1fk max %.1fk std %.1fk ; mean within-SM std %.1fk" % (
            len(per_sm), nper.min(), nper.max(), msm.min() / 1e3, msm.max() / 1e3, msm.std() / 1e3, ssm.mean() / 1e3))
        if E == 28:
            fin = full[:, :, 23]
            print("   finish cycles by lock-step group: warps 0-13: %.1fk ; warps 14-27: %.1fk" % (fin[:, :14].max(1).mean() / 1e3, fin[:, 14:].max(1).mean() / 1e3))
        bidx = full[:, 0, 26].astype(int)
        print("   CTA cycles by launch order: blockIdx < 148: %.1fk ; >= 148: %.1fk" % (cyc[bidx < 148].mean() / 1e3, cyc[bidx >= 148].mean() / 1e3))
        for k in (3, 4):
            sel = [np.mean(v) for v in per_sm.values() if len(v) == k]
            if sel: print("   SMs with %d CTAs: %d, mean CTA cycles %.1fk" % (k, len(sel), np.mean(sel) / 1e3))
        print("  per env: list entries/step mean %.0f max %.0f ; pairs mean %.0f max %.0f ; scanned rows mean %.1f p90 %.0f max %.0f" % (
            feat[:, 0].mean(), feat[:, 0].max(), feat[:, 1].mean(), feat[:, 1].max(), feat[:, 2].mean(), np.percentile(feat[:, 2], 90), feat[:, 2].max()))
        # per-env fit of the env's own finish time
        Xe = np.stack([np.ones(N), feat[:, 0], feat[:, 1], feat[:, 2]], 1)
        ce, *_ = np.linalg.lstsq(Xe, feat[:, 3], rcond=None)
        print("  per-env fit of finish time on [1, entries, pairs, rows]:", np.round(ce, 2).tolist())
    acc = o.stats[8:].cpu().numpy().reshape(N, 32)[:, 27:30].astype(np.float64).mean(0) / K
    print("  per env step (mean): wait at the pooled phases' opening barriers %.1fk, key-chain wait + key sides %.1fk ; per sub-step: front masks + list extension %.1fk" % (acc[0] * K / 1e3, acc[1] * K / 1e3, acc[2] / 1e3))
    pool = o.stats[8:].cpu().numpy().reshape(N, 32)[:, 30:32].astype(np.float64).mean(0)
    print("  pooled work per warp and env step (mean): scan rounds %.1fk, cell chunks %.1fk cycles" % tuple(pool / 1e3))
    extra = tr[:, 16:20].mean(0)
    tr = tr[:, :4 + 3 * K]
    mean = tr.mean(0); mx = tr.max(0)
    print("  abs stamps: e known %.1fk, copies issued %.1fk, prefetch_front done %.1fk, key chain done %.1fk" % tuple(extra / 1e3))
    d = np.diff(np.concatenate([[0], mean]))  # note: the key schedule sits between stamp 3 and "s0 owner-pre"
    print(mode, f"kernel {e0.elapsed_time(e1)*1e3:.1f} us = {e0.elapsed_time(e1)*1e3*1.965:.0f} kcycles*1e-3")
    print("  " + "  ".join(f"{n}:{x/1e3:.1f}k" for n, x in zip(names, d)))
    print("  last stamp mean %.1fk max %.1fk" % (mean[-1] / 1e3, mx[-1] / 1e3))
