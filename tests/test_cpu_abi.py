"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol the header declares,
the generated weight layout is current and consistent with the reference state_dict schema, argument validation
works without a device, and the host-side plumbing matches the oracle.  No compute call is made here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from trafficbots_b200 import _native
    _native.build()
    return _native.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "trafficbots_b200.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"tb_status", "tb_block", "tb_state_field"}
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), name
    from trafficbots_b200 import _native
    assert declared == set(_native.EXPORTS) | set(_native.training_signatures())  # tb_tr_*: argtypes parsed from the header


def test_weight_table_matches_state_dict_schema(lib):
    from trafficbots_b200 import weights
    spec = weights.state_dict_spec()
    n = lib.tb_weight_count()
    total = 0
    seen = set()
    for i in range(n):
        name = lib.tb_weight_name(i).decode()
        rows, cols = lib.tb_weight_rows(i), lib.tb_weight_cols(i)
        assert name in spec
        assert tuple(spec[name]) == ((rows,) if cols == 0 else (rows, cols)), name
        total += (rows + 3) // 4 * 4 if cols == 0 else (cols + 3) // 4 * 4 * rows
        seen.add(name)
    # [fp32 blob | pad to 1 KB | 64 KB tensor-core blocks]
    n_blk = sum((lib.tb_weight_rows(i) // 128) * (lib.tb_weight_cols(i) // 128) for i in range(n)
                if lib.tb_weight_cols(i) and lib.tb_weight_rows(i) % 128 == 0 and lib.tb_weight_cols(i) % 128 == 0)
    assert n_blk == lib.tb_tc_block_count()
    assert (total * 4 + 1023) // 1024 * 1024 + n_blk * 65536 == lib.tb_packed_weight_bytes()
    # every key of the reference state_dict is either packed, an alias of a packed tensor or a duplicated buffer
    for k in spec:
        if k in seen:
            continue
        assert weights._alias_of(k) in seen or k.startswith("pre_processing.latent.") or k.endswith("pl_node_ohe"), k
    assert lib.tb_weight_name(n) is None and lib.tb_weight_rows(-1) == -1


def test_generated_header_is_current():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_weight_layout
    assert open(gen_weight_layout.OUT).read() == gen_weight_layout.render()


def test_argument_validation_without_device(lib):
    from trafficbots_b200 import _native as nt
    good = nt.TbDims(2, 6, 64, 1024, 40, 11, 91, 90)
    assert lib.tb_rollout_state_bytes(C.byref(good)) > 0
    offs = [lib.tb_rollout_state_offset(C.byref(good), f) for f in range(9)]
    assert offs == sorted(offs) and offs[0] == 0 and all(o % 256 == 0 for o in offs)
    assert lib.tb_rollout_state_offset(C.byref(good), 9) == C.c_size_t(-1).value
    assert lib.tb_encode_workspace_bytes(C.byref(good)) >= 2 * 1024 * (128 + 256) * 4
    for bad in (nt.TbDims(0, 1, 8, 8, 8, 11, 91, 90), nt.TbDims(1, 1, 0, 8, 8, 11, 91, 90), nt.TbDims(70000, 1, 8, 8, 8, 11, 91, 90)):
        assert lib.tb_rollout_state_bytes(C.byref(bad)) == 0
        assert lib.tb_rollout_init(C.byref(bad), None, None, None, None) == -1
    assert lib.tb_rollout_init(C.byref(good), None, None, None, None) == -2
    assert lib.tb_encode_scene(C.byref(good), None, None, None, None, None) == -2
    assert lib.tb_pack_weights(None, None, None) == -2
    assert lib.tb_kv_project(7, 0, 16, 1, 16, 16, None) == -1
    assert lib.tb_kv_project(1, 1, 16, 1, 16, 16, None) == -1  # the global map layer has one layer only
    assert lib.tb_xlayer(2, 0, 16, 16, 4, 8, 16, 16, 8, 3, 0, 16, 16, None) == -1  # n_batch % kv_share


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from trafficbots_b200 import _native as nt, engine, weights
    with pytest.raises(nt.TbError):
        engine.Engine(weights.init_state_dict(0))
    with pytest.raises(nt.TbError):
        nt.dev_ptr(torch.zeros(4), "f32", name="x")


def test_teacher_forcing_mask_matches_oracle():
    import trafficbots_oracle as orc
    from trafficbots_b200 import host, synthetic
    batch = synthetic.make_batch(3, n_agent=16, n_pl=16, seed=5)
    v = batch["agent/valid"]
    for spawn, warm in ((10, 10), (90, 10), (0, -1), (5, 3)):
        assert torch.equal(host.teacher_forcing_mask(v, spawn, warm), orc.teacher_forcing_mask(v, spawn, warm))
