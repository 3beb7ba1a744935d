"""ctypes binding of `libtrafficbots_b200.so` (C ABI: `include/trafficbots_b200.h`).

The library is built in-tree by `build()` (nvcc, sm_100a only) and loaded from the package directory.  There is
no fallback: if the shared object is missing or a symbol is absent, importing the product path raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from typing import Optional

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("TB_LIB_PATH") or os.path.join(PKG_DIR, "libtrafficbots_b200.so")  # TB_LIB_PATH: A/B builds
CSRC = os.path.join(PKG_DIR, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC"]

TB_OK = 0
STATUS = {0: "TB_OK", -1: "TB_ERR_BAD_SHAPE", -2: "TB_ERR_NULL", -3: "TB_ERR_LAUNCH", -4: "TB_ERR_UNSUPPORTED",
          -5: "TB_ERR_ALIGN"}

# tb_block
BLOCK_MAP_DENSETNT, BLOCK_MAP_SELF_ATTN, BLOCK_AS2PL, BLOCK_AS2TL, BLOCK_INTERACTION, BLOCK_LATENT_PRIOR_INT, \
    BLOCK_LATENT_POST_INT = range(7)
# tb_gru / tb_mlp
GRU_POLICY, GRU_LATENT_PRIOR, GRU_LATENT_POST, GRU_DEST = range(4)
MLP_LATENT_PRIOR_MEAN, MLP_LATENT_POST_MEAN = range(2)
# tb_state_field
(STATE_AGENT_STATE, STATE_VALID, STATE_KILLED, STATE_VEL, STATE_ACC, STATE_YAW_RATE, STATE_GOAL_VALID, STATE_STICKY,
 STATE_HIDDEN) = range(9)

EXPORTS = (
    "tb_weight_count", "tb_weight_name", "tb_weight_rows", "tb_weight_cols", "tb_packed_weight_bytes",
    "tb_pack_weights", "tb_encode_workspace_bytes", "tb_encode_scene", "tb_kv_project", "tb_xlayer",
    "tb_rollout_state_bytes", "tb_rollout_state_offset", "tb_rollout_init", "tb_rollout_steps", "tb_step_front", "tb_step_back", "tb_rollout",
    "tb_launch_count", "tb_kv_tc_bytes", "tb_tc_block_count", "tb_tc_first_block", "tb_tc_selftest",
    "tb_gru_sequence", "tb_gru_workspace_bytes", "tb_mlp_head", "tb_dest_workspace_bytes", "tb_dest_logits", "tb_xlayer_tc",
    "tb_rule_workspace_bytes", "tb_rule_checks", "tb_post_process", "tb_womd_record_bytes", "tb_womd_pack",
)


class TbError(RuntimeError):
    pass


class TbDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_scene", "n_mode", "n_agent", "n_pl", "n_tl", "n_step_hist", "n_step_gt", "n_step", "n_cta_per_mode")]


def _ptr_struct(name, fields):
    return type(name, (C.Structure,), {"_fields_": [(f, C.c_void_p) for f in fields]})


SCENE_IN_FIELDS = ("map_valid", "map_type", "map_pos", "map_dir", "agent_valid", "agent_pos", "agent_yaw", "agent_vel",
                   "agent_spd", "agent_yaw_rate", "agent_acc", "agent_size", "agent_type", "tl_valid", "tl_state",
                   "tl_pos", "tl_dir")
SCENE_OUT_FIELDS = ("map_feature", "map_feature_valid", "agent_feature", "tl_feature", "kv_map", "kv_tl", "kv_map_tc", "kv_tl_tc",
                    "n_key_map", "n_key_tl")
ROLLOUT_IN_FIELDS = ("map_feature", "map_feature_valid", "kv_map", "kv_tl", "tl_valid", "gt_valid", "gt_pos", "gt_yaw",
                     "gt_spd", "gt_vel", "gt_acc", "gt_yaw_rate", "tf_mask", "agent_type", "agent_size", "map_boundary",
                     "map_valid", "map_type", "map_pos", "map_dir", "goal_gt", "latent_sample", "latent_logp", "dest",
                     "goal_valid", "kv_map_tc", "kv_tl_tc", "n_key_map", "n_key_tl")
ROLLOUT_OUT_FIELDS = ("preds", "valid", "override_masks", "diffbar_rewards", "diffbar_rewards_valid",
                      "action_log_probs", "latent_log_probs", "violations", "trace_policy_feature",
                      "trace_action_mean")
TbSceneIn = _ptr_struct("TbSceneIn", SCENE_IN_FIELDS)
TbSceneOut = _ptr_struct("TbSceneOut", SCENE_OUT_FIELDS)
TbRolloutIn = _ptr_struct("TbRolloutIn", ROLLOUT_IN_FIELDS)
TbRolloutOut = _ptr_struct("TbRolloutOut", ROLLOUT_OUT_FIELDS)


class TbRuleIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("preds", "valid", "override_masks", "outside_map_this_step", "gt_valid", "gt_pos", "gt_yaw", "gt_spd",
                 "agent_type", "agent_size", "map_valid", "map_type", "map_pos", "map_dir", "tl_valid", "tl_pos", "tl_state")] + \
               [("n_tl_frame", C.c_int32), ("enable_mask", C.c_int32), ("collision_size_scale", C.c_float),
                ("w_collision", C.c_float), ("reduce_collision_with_max", C.c_int32)]


class TbPostCfg(C.Structure):
    _fields_ = [("k_pred", C.c_int32), ("score_temperature", C.c_float), ("use_ade", C.c_int32), ("n_mtr", C.c_int32),
                ("mtr_nms_thresh", C.c_float * 3), ("n_mpa", C.c_int32), ("mpa_nms_thresh", C.c_float * 3)]


class TbWomdIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("agent_role", "agent_valid", "agent_pos", "agent_size", "agent_yaw", "agent_vel",
                                          "agent_type", "waymo_trajs", "waymo_scores")] + \
               [(n, C.c_int32) for n in ("n_agent", "n_pred", "n_step_future", "n_step_gt_frames", "step_gt", "step_current", "m_joint")]


class TbWomdOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("prediction_trajectory", "prediction_score", "ground_truth_trajectory",
                                          "ground_truth_is_valid", "prediction_ground_truth_indices_mask", "object_type")] + \
               [("scene_stride_bytes", C.c_int64), ("overflow", C.c_void_p)]


TbRuleOut = _ptr_struct("TbRuleOut", ("violations", "diffbar_rewards", "diffbar_rewards_valid"))
RULE_BITS = {"collided": 1, "run_road_edge": 2, "run_red_light": 4, "passive": 8}
OPT_VIOLATION_KEYS = ("collided", "collided_this_step", "run_road_edge", "run_road_edge_this_step", "run_red_light",
                      "run_red_light_this_step", "passive", "passive_this_step")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a (works without a GPU).  Output: trafficbots_b200/libtrafficbots_b200.so.
    Every translation unit is compiled to `build/<name>.o` (in parallel, only when older than the sources / headers or
    when the flags changed) and the objects are linked into the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    import concurrent.futures as cf
    import hashlib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("TB_NVCC_FLAGS", "").split()  # e.g. -DTB_TRACE_DETAIL (tools/trace_rollout.py)
    cflags = [f for f in NVCC_FLAGS if f != "--shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    tag = hashlib.sha1((" ".join(cflags) + "|" + os.path.basename(LIB_PATH)).encode()).hexdigest()[:10]
    obj_dir = os.path.join(PKG_DIR, "build", tag)
    os.makedirs(obj_dir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    t_hdr = max(os.path.getmtime(h) for h in headers)

    def compile_one(src: str):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(t_hdr, os.path.getmtime(src)):
            return obj, ""
        cmd = [nvcc] + cflags + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise TbError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        done = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + [o for o, _ in done]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise TbError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print("".join(log for _, log in done))
    return LIB_PATH


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded library; raises TbError if it has not been built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(the CUDA extension is the only implementation of this path)")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise TbError(f"{LIB_PATH} does not export {name}")
    L.tb_weight_count.restype = C.c_int32
    L.tb_weight_name.restype = C.c_char_p
    L.tb_weight_name.argtypes = [C.c_int32]
    L.tb_weight_rows.restype = C.c_int32
    L.tb_weight_rows.argtypes = [C.c_int32]
    L.tb_weight_cols.restype = C.c_int32
    L.tb_weight_cols.argtypes = [C.c_int32]
    L.tb_packed_weight_bytes.restype = C.c_size_t
    L.tb_pack_weights.restype = C.c_int32
    L.tb_pack_weights.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    L.tb_encode_workspace_bytes.restype = C.c_size_t
    L.tb_encode_workspace_bytes.argtypes = [C.POINTER(TbDims)]
    L.tb_encode_scene.restype = C.c_int32
    L.tb_encode_scene.argtypes = [C.POINTER(TbDims), C.POINTER(TbSceneIn), C.c_void_p, C.POINTER(TbSceneOut),
                                  C.c_void_p, C.c_void_p]
    L.tb_kv_project.restype = C.c_int32
    L.tb_kv_project.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_xlayer.restype = C.c_int32
    L.tb_xlayer.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                            C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_rollout_state_bytes.restype = C.c_size_t
    L.tb_rollout_state_bytes.argtypes = [C.POINTER(TbDims)]
    L.tb_rollout_state_offset.restype = C.c_size_t
    L.tb_rollout_state_offset.argtypes = [C.POINTER(TbDims), C.c_int32]
    L.tb_rollout_init.restype = C.c_int32
    L.tb_rollout_init.argtypes = [C.POINTER(TbDims), C.POINTER(TbRolloutIn), C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_rollout_steps.restype = C.c_int32
    L.tb_rollout_steps.argtypes = [C.POINTER(TbDims), C.POINTER(TbRolloutIn), C.c_void_p, C.c_void_p,
                                   C.POINTER(TbRolloutOut), C.c_int32, C.c_int32, C.c_void_p]
    L.tb_step_front.restype = C.c_int32
    L.tb_step_front.argtypes = [C.POINTER(TbDims), C.POINTER(TbRolloutIn), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.tb_step_back.restype = C.c_int32
    L.tb_step_back.argtypes = [C.POINTER(TbDims), C.POINTER(TbRolloutIn), C.c_void_p, C.c_void_p,
                               C.POINTER(TbRolloutOut), C.c_int32, C.c_void_p]
    L.tb_rollout.restype = C.c_int32
    L.tb_rollout.argtypes = [C.POINTER(TbDims), C.POINTER(TbRolloutIn), C.c_void_p, C.c_void_p,
                             C.POINTER(TbRolloutOut), C.c_void_p]
    L.tb_launch_count.restype = C.c_int64
    L.tb_kv_tc_bytes.restype = C.c_size_t
    L.tb_kv_tc_bytes.argtypes = [C.POINTER(TbDims), C.c_int32]
    L.tb_tc_block_count.restype = C.c_int32
    L.tb_tc_first_block.restype = C.c_int32
    L.tb_tc_first_block.argtypes = [C.c_int32]
    L.tb_tc_selftest.restype = C.c_int32
    L.tb_tc_selftest.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.tb_xlayer_tc.restype = C.c_int32
    L.tb_xlayer_tc.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_gru_sequence.restype = C.c_int32
    L.tb_gru_sequence.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_gru_workspace_bytes.restype = C.c_size_t
    L.tb_gru_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    L.tb_mlp_head.restype = C.c_int32
    L.tb_mlp_head.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.tb_dest_workspace_bytes.restype = C.c_size_t
    L.tb_dest_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.tb_dest_logits.restype = C.c_int32
    L.tb_dest_logits.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 11
    L.tb_rule_workspace_bytes.restype = C.c_size_t
    L.tb_rule_workspace_bytes.argtypes = [C.POINTER(TbDims)]
    L.tb_rule_checks.restype = C.c_int32
    L.tb_rule_checks.argtypes = [C.POINTER(TbDims), C.POINTER(TbRuleIn), C.POINTER(TbRuleOut), C.c_void_p, C.c_void_p]
    L.tb_post_process.restype = C.c_int32
    L.tb_post_process.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.POINTER(TbPostCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    L.tb_womd_record_bytes.restype = C.c_size_t
    L.tb_womd_record_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
    L.tb_womd_pack.restype = C.c_int32
    L.tb_womd_pack.argtypes = [C.c_int32, C.POINTER(TbWomdIn), C.POINTER(TbWomdOut), C.c_void_p]
    for name, argtypes in training_signatures().items():  # tb_tr_*: argtypes parsed from the header (single source)
        if not hasattr(L, name):
            raise TbError(f"{LIB_PATH} does not export {name}")
        fn = getattr(L, name)
        fn.restype = C.c_int32
        fn.argtypes = argtypes
    _lib = L
    return L


_CTYPE = {"int64_t": C.c_int64, "int32_t": C.c_int32, "uint32_t": C.c_uint32, "float": C.c_float}


def training_signatures():
    """{name: ctypes argtypes} of every `tb_tr_*` entry point declared in include/trafficbots_b200.h."""
    import re
    text = open(os.path.join(ROOT, "include", "trafficbots_b200.h")).read()
    out = {}
    for m in re.finditer(r"int32_t (tb_tr_[a-z0-9_]+)\(([^)]*)\);", text):
        args = []
        for a in m.group(2).split(","):
            a = a.strip()
            if "*" in a:
                args.append(C.c_void_p)
            else:
                args.append(_CTYPE[a.replace("const ", "").split()[0]])
        out[m.group(1)] = args
    return out


def check(rc: int, what: str) -> None:
    if rc != TB_OK:
        raise TbError(f"{what} failed: {STATUS.get(rc, rc)}")


_DT = {"f32": torch.float32, "u8": (torch.bool, torch.uint8), "i64": torch.int64, "i32": torch.int32}


def dev_ptr(t: Optional[torch.Tensor], kind: str, shape=None, name: str = "tensor", optional: bool = False) -> Optional[int]:
    """device pointer of a dense CUDA tensor after checking dtype / shape / layout (None -> NULL if optional)."""
    if t is None:
        if optional:
            return None
        raise TbError(f"{name}: required tensor is None")
    want = _DT[kind]
    ok = t.dtype in want if isinstance(want, tuple) else t.dtype == want
    if not ok:
        raise TbError(f"{name}: dtype {t.dtype}, expected {kind}")
    if not t.is_cuda:
        raise TbError(f"{name}: expected a CUDA tensor (the hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise TbError(f"{name}: tensor must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise TbError(f"{name}: shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t.data_ptr()


def current_stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
