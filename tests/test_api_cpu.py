"""Host-side logic that must work on a CPU box: the drop-in API surface, checkpoint round trip, C-ABI symbols."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import rectified_flow_vision_b200 as pkg
from rectified_flow_vision_b200 import engine
from tests import util


def test_public_names_match_reference():
    assert set(pkg.__all__) == {'UNet', 'count_parameters', 'BaseFlowModel', 'train_base_flow', 'RectifiedFlowModel',
                                'generate_reflow_pairs', 'train_rectified_flow', 'iterative_reflow'}


def test_signatures_match_reference():
    sig = inspect.signature
    assert list(sig(pkg.BaseFlowModel.__init__).parameters)[1:] == [
        'image_size', 'in_channels', 'model_channels', 'channel_mult', 'num_res_blocks', 'attention_resolutions',
        'dropout', 'device']
    assert list(sig(pkg.BaseFlowModel.sample).parameters)[1:] == ['noise', 'num_steps', 'batch_size', 'return_trajectory']
    assert sig(pkg.BaseFlowModel.sample).parameters['num_steps'].default == 100
    assert list(sig(pkg.BaseFlowModel.sample_with_trajectory).parameters)[1:] == ['noise', 'num_steps', 'save_every']
    assert list(sig(pkg.RectifiedFlowModel.compute_straightness).parameters)[1:] == ['x0', 'x1', 'num_points']
    assert list(sig(pkg.generate_reflow_pairs).parameters)[:4] == ['teacher_model', 'num_pairs', 'batch_size', 'num_steps']
    assert sig(pkg.generate_reflow_pairs).parameters['batch_size'].default == 32
    assert list(sig(pkg.train_rectified_flow).parameters) == ['model', 'x0_data', 'x1_data', 'epochs', 'batch_size', 'lr',
                                                              'save_path', 'save_every']


def test_state_dict_layout_default():
    m = pkg.BaseFlowModel(device="cpu")
    sd = m.state_dict()
    assert len(sd) == 174 and pkg.count_parameters(m) == 11255363
    assert all(k.startswith("velocity_net.") for k in sd) and all(v.dtype == torch.float32 for v in sd.values())
    assert sd["velocity_net.enc_blocks.2.shortcut.weight"].shape == (128, 64, 1, 1)
    assert sd["velocity_net.dec_blocks.0.conv1.weight"].shape == (256, 512, 3, 3)
    assert sd["velocity_net.upsamples.1.1.weight"].shape == (128, 128, 3, 3)
    assert sd["velocity_net.mid_attn.qkv.weight"].shape == (768, 256, 1, 1)
    assert sd["velocity_net.output_conv.2.weight"].shape == (3, 64, 3, 3)
    assert not any(k.startswith("velocity_net.downsamples.2") or k.startswith("velocity_net.upsamples.2") for k in sd)


def test_checkpoint_roundtrip_reference_format(tmp_path):
    torch.manual_seed(3)
    m = pkg.RectifiedFlowModel(device="cpu")
    path = str(tmp_path / "ck" / "model_final.pt")
    m.save(path)
    blob = torch.load(path)
    assert set(blob.keys()) == {"state_dict", "config"} and blob["config"] == {"image_size": 64, "in_channels": 3}
    m2 = pkg.BaseFlowModel(device="cpu")
    m2.load(path)
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])


def test_from_base_model_uses_default_architecture():
    base = pkg.BaseFlowModel(image_size=32, model_channels=128, channel_mult=[1, 2], device="cpu")
    r = pkg.RectifiedFlowModel.from_base_model(base)
    assert r.image_size == 32 and r.velocity_net.model_channels == 64 and r.velocity_net.channel_mult == [1, 2, 4]
    assert r.reflow_iteration == 0


def test_interpolation_identities():
    m = pkg.BaseFlowModel(image_size=32, channel_mult=[1], num_res_blocks=1, device="cpu")
    x0, x1 = torch.randn(2, 3, 8, 8), torch.randn(2, 3, 8, 8)
    for tv, want in ((0.0, x0), (1.0, x1), (0.5, (x0 + x1) / 2)):
        xt, tgt = m.get_interpolation(x0, x1, torch.full((2,), tv))
        assert torch.allclose(xt, want, atol=1e-6) and torch.equal(tgt, x1 - x0)


def test_no_cpu_fallback():
    m = pkg.BaseFlowModel(image_size=32, channel_mult=[1, 2], num_res_blocks=1, device="cpu")
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 32, 32), torch.zeros(1))
    with pytest.raises(RuntimeError):
        m.sample(batch_size=1, num_steps=1)
    with pytest.raises(RuntimeError, match="no CPU path"):   # the trainer is native too: no autograd fallback
        pkg.train_rectified_flow(m, torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32), epochs=1)


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(util.ROOT, "include", "rfv.h")).read()
    declared = set(re.findall(r"\b(rfv_[a-z_0-9]+)\s*\(", header))
    assert declared == set(engine.SYMBOLS), declared ^ set(engine.SYMBOLS)
    lib = engine.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rfv_abi_version() == 1
    # argument checking works without a GPU (no compute call is made)
    assert lib.rfv_create(None, None) != 0
    assert b"null" in lib.rfv_last_error()


def test_benchmark_csv_schema_matches_reference(tmp_path):
    """experiments/benchmark.py:252-262 writes num_steps,base_time_ms,rect_time_ms,base_img_per_sec,rect_img_per_sec,speedup."""
    from rectified_flow_vision_b200 import benchmark as B

    class Stub:
        def eval(self):
            return self

        def sample(self, noise=None, num_steps=1):
            return noise

    br = B.benchmark_speed(Stub(), 8, [1, 2], 8, "cpu", num_runs=2, batch_size=4)
    assert [set(r) for r in br] == [{'num_steps', 'total_time', 'time_per_image', 'images_per_second', 'time_std', 'num_samples'}] * 2
    out = tmp_path / "results" / "benchmark_results.csv"
    B.write_results_csv(str(out), br, br)
    lines = out.read_text().strip().splitlines()
    assert lines[0] == "num_steps,base_time_ms,rect_time_ms,base_img_per_sec,rect_img_per_sec,speedup"
    assert len(lines) == 3 and lines[1].startswith("1,") and lines[1].endswith(",1.0")


def test_engine_flags_are_distinct_bits():
    """include/rfv.h: every RFV_FLAG_* is its own bit, clear of the cluster-size field (bits 8-10) -- two switches sharing a bit
    would silently change which kernel an A/B test measures."""
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "rfv.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(RFV_FLAG_\w+)\s+(\d+)", hdr)}
    assert len(flags) >= 12
    seen = 0
    for name, v in flags.items():
        assert v > 0 and v & (v - 1) == 0, (name, v)
        assert not (v & seen), name
        assert not (v & (7 << 8)), name
        seen |= v
    from rectified_flow_vision_b200 import engine as E
    assert E.FLAG_TRAIN == flags["RFV_FLAG_TRAIN"] and E.FLAG_ONE_LANE == flags["RFV_FLAG_ONE_LANE"]
    assert E.FLAG_NO_UMMA == flags["RFV_FLAG_NO_UMMA"] and E.FLAG_KEEP_ACTS == flags["RFV_FLAG_KEEP_ACTS"]


def test_benchmark_report_text_matches_reference(tmp_path):
    """utils/visualization.py:210-253: the text of benchmark_report.txt, byte for byte against the file the reference's own
    create_summary_report wrote for the same results dictionary (oracle/make_golden_report.py)."""
    import json
    from rectified_flow_vision_b200 import benchmark as B
    results = json.load(open(os.path.join(util.GOLD, "benchmark_report.json")))
    want = open(os.path.join(util.GOLD, "benchmark_report.txt")).read()
    assert B.summary_report_text(results) == want
    path = B.create_summary_report(results, str(tmp_path / "results"))
    assert os.path.basename(path) == "benchmark_report.txt" and open(path).read() == want
    zero = {"base_model": [dict(results["base_model"][0])], "rectified_model": [dict(results["rectified_model"][0], time_per_image=0.0)]}
    txt = B.summary_report_text(zero)          # a zero rectified time prints a 0.00x row and no conclusions (reference :236, :246)
    assert "0.00      x" in txt and "Average speedup" not in txt
