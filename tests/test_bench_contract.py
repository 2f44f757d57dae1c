"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port on the host cores) prints exactly
one JSON line on stdout with the keys the driver reads; the GPU arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

import common

BENCH = os.path.join(common.ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=common.ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("env steps/sec") and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_gpu_arm_fails_loudly_without_cuda():
    p = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=common.ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_stdout_carries_only_the_json_line():
    """The GPU arm sends library chatter on fd 1 (NCCL banner) to stderr and writes its JSON line to the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench\n"
            "fd = bench.redirect_stdout_to_stderr()\n"
            "print('python chatter'); os.system('echo child chatter')\n"
            "bench.emit_json({'metric': 'm', 'value': 1.5}, fd); bench.restore_stdout(fd); print('after')\n") % common.ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert p.stdout.splitlines() == ['{"metric": "m", "value": 1.5}', "after"]
    assert "python chatter" in p.stderr and "child chatter" in p.stderr


def test_config_flags_select_the_baseline_configurations(monkeypatch):
    """--config 2..5 = BASELINE.json configs[1..4] (envs per GPU, clips, seed); --scaling strong keeps the global batch."""
    sys.path.insert(0, common.ROOT)
    import bench

    def parse(argv, world=1):
        monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
        monkeypatch.setenv("WORLD_SIZE", str(world))
        return bench.parse()

    a = parse([])
    assert (a.config, a.envs_per_gpu, a.motions, a.seed, a.scaling) == (2, 4096, ["walk"], 0, "weak")
    a = parse(["--config", "3"])
    assert (a.envs_per_gpu, a.motions, a.seed) == (16384, ["spinkick"], 1)
    a = parse(["--config", "4", "--gpus", "8"], world=8)
    assert (a.envs_per_gpu, a.envs_global, a.motions) == (8192, 65536, ["walk"])
    a = parse(["--config", "4", "--scaling", "strong", "--envs-global", "65536"], world=2)
    assert (a.envs_per_gpu, a.envs_global) == (32768, 65536)
    a = parse(["--config", "5"], world=8)
    assert (a.envs_per_gpu, a.envs_global, a.motions, a.seed) == (4096, 32768, ["walk", "dance_b", "spinkick"], 2)
    a = parse(["--motions", "run,walk", "--envs-per-gpu", "128"])
    assert a.motions == ["run", "walk"] and a.envs_per_gpu == 128 and a.envs_global == 128
    with pytest.raises(SystemExit):
        parse(["--scaling", "strong", "--envs-global", "100"], world=8)
