"""bench.py's reference arm runs without a GPU (it is the CPU path) and prints the JSON line the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"]


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, SC_BENCH_CPU_BUDGET_S="6")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    for k in REQUIRED:
        assert k in out, k
    assert out["impl"] == "reference" and out["n_gpus"] == 1 and out["steps"] == 2
    assert out["value"] > 0 and out["higher_is_better"] is True and out["unit"] == "samples/s"
    assert "workload" in out["config"] and "model" not in out["config"]
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_is_rank0_only_under_torchrun():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29512")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_torch_cufft_denominator_is_the_reference_op_sequence():
    """bench.py's PyTorch+cuFFT denominator restates the reference forward (it must not import oracle/ on the timed path);
    on CPU it has to agree bit-exactly with the oracle, which is pinned to the unmodified reference."""
    import torch
    import bench
    from oracle import spectral_conv_oracle as O
    for (B, Ci, Co, grid, modes) in [(2, 3, 4, (16, 12), (8, 6)), (2, 3, 3, (32,), (8,)), (1, 2, 3, (8, 6, 10), (4, 4, 6)),
                                     (2, 3, 4, (9, 11), (4, 5))]:
        x, w, bias, gy = O.make_inputs(B, Ci, Co, grid, modes, seed=1)
        y_ref = O.spectral_conv_forward(x, w, bias, modes)
        y = bench.torch_cufft_forward(x, w.tensor, bias, O.stored_n_modes(modes))
        assert torch.equal(y, y_ref)


def test_torch_layer_denominator_is_the_reference_fourier_layer():
    """bench.py's PyTorch denominator for the Fourier layer (f1 / f2) restates fno_block.py:377-414; on CPU it agrees with the block
    oracle, which is pinned to the unmodified reference FNOBlocks (tests/test_block_oracle.py)."""
    import torch
    import bench
    from conftest import block_oracle_kwargs, load_block_golden
    from oracle import fno_block_oracle as BO
    from oracle import spectral_conv_oracle as O
    for name, last in (("block_d2_default_mid", False), ("block_d2_default_last", True)):
        meta, io, params, _ = load_block_golden(name)
        i = meta["index"]
        prm = {"w": params[f"convs.{i}.weight.tensor"], "b": params[f"convs.{i}.bias"], "w_skip": params[f"fno_skips.{i}.conv.weight"],
               "w1": params[f"channel_mlp.{i}.fcs.0.weight"], "b1": params[f"channel_mlp.{i}.fcs.0.bias"],
               "w2": params[f"channel_mlp.{i}.fcs.1.weight"], "b2": params[f"channel_mlp.{i}.fcs.1.bias"],
               "gate": params[f"channel_mlp_skips.{i}.weight"]}
        y = bench.torch_layer_forward(io["x"], prm, O.stored_n_modes(meta["n_modes"]), last=last)
        y_ref = BO.fno_block_forward(io["x"], params, i, **block_oracle_kwargs(meta))
        assert (y - y_ref).abs().max() <= 1e-6 * y_ref.abs().max()
        assert (y - io["y"]).abs().max() <= 2e-5 * io["y"].abs().max()
