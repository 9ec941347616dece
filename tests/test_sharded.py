"""Row sharding: world_size-2 runs of tests/sharded_worker.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(mode, *extra, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py"), mode, *extra]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_collectives_and_sharding_arithmetic_gloo():
    r = _launch("gloo")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 gloo ok" in r.stdout and "rank 1 gloo ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["BayesR", "BayesCpi", "BayesB"])   # BayesB (BASELINE config 3) added after the last 2-GPU run
def test_two_gpus_match_the_oracle(model):
    import hibayes_b200 as hb
    if hb.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = _launch("gpu", model, port=29612)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 gpu %s ok" % model in r.stdout and "rank 1 gpu %s ok" % model in r.stdout


@pytest.mark.gpu
def test_two_gpus_covariates_random_effects_single_step():
    import hibayes_b200 as hb
    if hb.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = _launch("gpu_fx", port=29613)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 gpu_fx ok" in r.stdout and "rank 1 gpu_fx ok" in r.stdout
