"""Generates tests/golden/schedule_kat.json: known answers for the integer / float32 schedules.

The reference ships no golden vectors and cannot run here (diffusers absent), so these are
re-derived INDEPENDENTLY of oracle/ and of the product: plain numpy float32 arithmetic written
from the published formulas, then cross-checked against the constants
the survey lists in appendix A.6 (hard-coded below).  Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

f32 = np.float32


def linspace_f32(start, end, n):
    """torch.linspace(float32): symmetric evaluation from both ends."""
    start, end = f32(start), f32(end)
    step = (end - start) / f32(n - 1)
    i = np.arange(n)
    half = n // 2
    lo = start + step * i.astype(f32)
    hi = end - step * (n - 1 - i).astype(f32)
    return np.where(i < half, lo, hi).astype(f32)


def alphas_cumprod():
    betas = linspace_f32(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
    alphas = (f32(1.0) - betas).astype(f32)
    out = np.empty(1000, dtype=f32)
    acc = f32(1.0)
    for i, a in enumerate(alphas):
        acc = f32(acc * a)
        out[i] = acc
    return out


def ddim_timesteps(n):
    return ((np.arange(n) * (1000 // n)).round()[::-1].astype(np.int64) + 1).tolist()


def dpm_timesteps(n):
    return ((np.arange(n + 1) * (1000 // (n + 1))).round()[::-1][:-1].astype(np.int64) + 1).tolist()


def dpm_sigmas(n, ac, final="zero"):
    full = (((f32(1) - ac) / ac) ** f32(0.5)).astype(f32)
    ts = np.array(dpm_timesteps(n))
    sig = np.interp(ts, np.arange(1000), full)
    last = 0.0 if final == "zero" else float(full[0])
    return np.concatenate([sig, [last]]).astype(f32)


def lcm_timesteps(n):
    origin = (np.arange(1, 51) * 20 - 1)[::-1]
    idx = np.floor(np.linspace(0, 50, num=n, endpoint=False)).astype(np.int64)
    return origin[idx].tolist()


def pndm_timesteps(n):
    t = (np.arange(n) * (1000 // n)).round() + 1
    return np.concatenate([t[:-1], t[-2:-1], t[-1:]])[::-1].astype(np.int64).tolist()


def two_scheduler(n1, k):
    first_full = ddim_timesteps(n1)
    first = first_full[:k]
    dist = [abs(t - first[-1]) for t in first_full]
    return first, first_full[int(np.argmin(dist)):]


def temb(t, idx):
    j = np.arange(160, dtype=f32)
    freq = np.exp((f32(-np.log(10000.0)) * j / f32(160)).astype(f32)).astype(f32)
    ang = (f32(t) * freq).astype(f32)
    emb = np.concatenate([np.cos(ang), np.sin(ang)]).astype(f32)
    return [float(emb[i]) for i in idx]


def main():
    ac = alphas_cumprod()
    kat = {
        "alphas_cumprod": {str(i): float(ac[i]) for i in (0, 1, 500, 999)},
        "ddim_20": ddim_timesteps(20), "ddim_50": ddim_timesteps(50),
        "dpm_25": dpm_timesteps(25),
        "dpm_25_sigmas_head": [float(x) for x in dpm_sigmas(25, ac)[:3]],
        "dpm_25_sigmas_tail": [float(x) for x in dpm_sigmas(25, ac)[-3:]],
        "sigma_min": float((((f32(1) - ac[0]) / ac[0]) ** f32(0.5))),
        "lcm_4": lcm_timesteps(4), "lcm_2": lcm_timesteps(2), "lcm_1": lcm_timesteps(1),
        "pndm_50": pndm_timesteps(50),
        "two_10_3": two_scheduler(10, 3), "two_20_10": two_scheduler(20, 10),
        "ddim_951": {"a_t": float(ac[951]), "a_prev": float(ac[901])},
        "ddim_1": {"a_t": float(ac[1]), "a_prev": float(ac[0])},
        "lcm_c_skip_999": 0.25 / ((999 * 10.0) ** 2 + 0.25), "lcm_c_skip_259": 0.25 / ((259 * 10.0) ** 2 + 0.25),
        "temb_951": temb(951, (0, 1, 160, 161)),
        "unet_params": 859520964, "flop_per_sample": 803.27e9, "deepcache_branch0_flop": 63.25e9,
    }
    # constants printed in SURVEY.md appendix A.6 -- the derivation above must reproduce them
    assert abs(kat["alphas_cumprod"]["0"] - 0.9991499781608582) < 1e-9
    assert abs(kat["alphas_cumprod"]["1"] - 0.9982960224151611) < 1e-9
    assert abs(kat["alphas_cumprod"]["500"] - 0.27633246779441833) < 1e-7
    # 1000 sequential float32 products: numpy and torch may differ in the last ulps
    assert abs(kat["alphas_cumprod"]["999"] / 0.00466009508818388 - 1) < 1e-5
    assert kat["ddim_20"][:3] == [951, 901, 851] and kat["ddim_20"][-2:] == [51, 1]
    assert kat["dpm_25"] == [951, 913, 875, 837, 799, 761, 723, 685, 647, 609, 571, 533, 495, 457, 419, 381, 343,
                             305, 267, 229, 191, 153, 115, 77, 39]
    assert np.allclose(kat["dpm_25_sigmas_head"], [11.028335571, 8.943601608, 7.334478855], rtol=2e-5, atol=1e-9)
    assert np.allclose(kat["dpm_25_sigmas_tail"], [0.291284561, 0.196303144, 0.0], rtol=2e-5, atol=1e-9)
    assert abs(kat["sigma_min"] / 0.029167533 - 1) < 2e-5
    assert kat["lcm_4"] == [999, 759, 499, 259] and kat["lcm_2"] == [999, 499] and kat["lcm_1"] == [999]
    assert kat["pndm_50"][:5] == [981, 961, 961, 941, 921] and len(kat["pndm_50"]) == 51
    assert kat["two_10_3"] == ([901, 801, 701], [701, 601, 501, 401, 301, 201, 101, 1])
    assert kat["two_20_10"][0][-1] == 501 and kat["two_20_10"][1][:2] == [501, 451]
    assert abs(kat["ddim_951"]["a_t"] - 0.00815500) < 1e-7 and abs(kat["ddim_951"]["a_prev"] - 0.01400489) < 1e-7
    assert np.allclose(kat["temb_951"], [-0.6195915937, 0.7689372897, 0.7849243283, -0.6393241882], atol=2e-4)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "schedule_kat.json")
    with open(out, "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
