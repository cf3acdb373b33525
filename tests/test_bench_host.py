"""Host-side pieces of bench.py that need no GPU: the algorithmic-byte model, the latency bound, and the reference arm
(`--impl reference`: whole meta-steps of the reference model timed on the host cores, one JSON line with the contract's keys)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_algorithmic_bytes_and_latency_bound_of_the_recurrent_kernels():
    import bench
    import msa_tts_b200 as pkg
    cfg = pkg.default_params()
    for k in ("enc_lstm_fwd", "attn_chain_fwd", "dec_lstm_fwd", "dec_lstm_bwd", "attn_chain_bwd", "enc_lstm_bwd"):
        ab = bench.algo_bytes(cfg, k)
        assert 1e6 < ab < 2e8, (k, ab)
        lb = bench.latency_bound(k, 1.0)
        assert lb["steps"] == (bench.L if k.startswith("enc_") else bench.T)
        assert abs(lb["bound_ms"] - lb["steps"] * lb["hand_offs_per_step"] * lb["hand_off_us"] * 1e-3) < 1e-12
        assert 0 < lb["frac"] == lb["bound_ms"] / 1.0
    # per-task-weight launches ("_pt", three tasks = 12 rows): the single-launch bytes at 12 rows + two more recurrent weight matrices
    from msa_tts_b200.config import rnn_dims
    Ha, A = rnn_dims(cfg)[0], cfg["attention_params"]["attention_dim"]
    for k in ("attn_chain_fwd", "attn_chain_bwd"):
        assert bench.algo_bytes(cfg, k + "_pt", 12) == bench.algo_bytes(cfg, k, 12) + 4.0 * 2 * (4 * Ha * Ha + A * Ha)
        assert bench.latency_bound(k + "_pt", 2.0)["bound_ms"] == bench.latency_bound(k, 2.0)["bound_ms"] == bench.latency_bound(k + "_grp", 2.0)["bound_ms"]
    # 88.8 MB forward / 114.6 MB backward at the bench shape
    assert abs(bench.algo_bytes(cfg, "attn_chain_fwd") - 88.8e6) < 0.1e6
    assert abs(bench.algo_bytes(cfg, "attn_chain_bwd") - 114.6e6) < 0.1e6
    p = bench.trainer_params(1)
    assert p["n_inner_train"] == 1 and p["meta_batch_size"] == 8 and p["track_higher_grads"] is False


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "meta_steps_per_s" and line["unit"] == "meta-steps/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0
    assert line["config"]["workload"].startswith("fomaml_meta_step_8tasks")
    # the unmodified reference model when baseline/_ref (or /root/reference) is importable, the oracle port otherwise
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["steps_requested"] == 1 and abs(line["value"] - 1000.0 / line["ms_per_step"]) < 1e-9
    assert line["e2e"] == {"value": line["value"], "unit": "meta-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0 < line["value"] < 10


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
