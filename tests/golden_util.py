"""Helpers shared by the golden-vector tests: load a fixture and regenerate its seeded inputs / parameters."""
import os

import numpy as np
import torch

from trafficbots_b200 import synthetic, weights

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("cfg1_s1_a8_p64_k1", "s3_a8_p64_k2", "s1_a64_p1024_k1")
# reference run with the four optional rule checks and the collision reward switched on (oracle/make_golden.py RULE_CASES)
RULE_CASES = ("s2_a16_p96_k2_rules", "s2_a12_p64_k1_rules_sum")
RULES_ON = {"collided": True, "run_road_edge": True, "run_red_light": True, "passive": True}
RULE_KEYS = tuple(k + s for k in ("collided", "run_road_edge", "run_red_light", "passive") for s in ("", "_this_step"))


def checksum(tensors) -> float:
    acc = 0.0
    for i, (k, v) in enumerate(sorted(tensors.items())):
        acc += (i + 1) * float(v.double().sum()) + 1e-3 * float(v.double().abs().sum())
    return acc


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    gold = {k.replace("__", "/"): torch.from_numpy(z[k]) for k in z.files if not k.startswith("meta__")}
    S, A, P, K, seed, wseed, sseed = [int(x) for x in z["meta__case"]]
    sd = weights.init_state_dict(wseed)
    meta = dict(S=S, A=A, P=P, K=K, seed=seed, wseed=wseed, sseed=sseed)
    if "meta__rules" in z.files:
        area, wcol, rmax = [float(x) for x in z["meta__rules"]]
        meta.update(area_scale=area, w_collision=wcol, reduce_with_max=bool(rmax))
        batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed, special_scenes=False, area_scale=area, plant_red_light=True)
    else:
        batch = synthetic.make_batch(S, n_agent=A, n_pl=P, seed=seed)
    # the regenerated inputs must be the ones the reference saw
    assert abs(checksum(batch) - float(z["meta__checksum_batch"])) <= 1e-6 * abs(float(z["meta__checksum_batch"]))
    assert abs(checksum(sd) - float(z["meta__checksum_weights"])) <= 1e-6 * abs(float(z["meta__checksum_weights"]))
    return gold, sd, batch, meta
