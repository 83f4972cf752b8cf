"""N>1 path: SFC-range partition + ghost cells + halo lists.  CPU: world_size 2 and 3 over gloo with the oracle standing in
for the device; GPU: the real thing over NCCL (needs >= 2 GPUs).  Acceptance: owned cells are
bit-identical to the single-domain run."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, mode, shape, ndist, steps=12, timeout=600, bc="walls", extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(HERE, "dist_worker.py"), "--mode", mode, "--shape", shape,
           "--ndist", str(ndist), "--steps", str(steps), "--bc", bc, *extra]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "PARTITION_PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,shape,ndist", [(2, "16,12,10", 19), (3, "12,12,8", 27)])
def test_partitioned_oracle_matches_single_domain_gloo(world, shape, ndist):
    launch(world, "oracle", shape, ndist)


@pytest.mark.parametrize("world,shape,ndist", [(2, "18,16,16", 19), (3, "10,10,10", 27), (2, "34,64", 9)])
def test_partitioned_pressure_boundary_velocity_halo_gloo(world, shape, ndist):
    """Pressure in-/outlet whose inward neighbours lie across the cut: the velocity halo."""
    launch(world, "oracle", shape, ndist, bc="pressure")


@pytest.mark.parametrize("world,case,collision", [(4, "sphere3d", "mrt"), (3, "step3d", "trt")])
def test_partitioned_baseline_configs_gloo(world, case, collision):
    """BASELINE.json configs[3] / [4] (3D sphere D3Q27 MRT, 3D step D3Q19 TRT with pressure outflow) cut into SFC ranges."""
    launch(world, "oracle", "0,0,0", 0, steps=8, extra=("--case", case, "--level", "5", "--collision", collision))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ndist", [("32,32,32", 19), ("24,16,16", 27), ("64,64", 9)])
def test_partitioned_gpu_matches_single_domain_nccl(shape, ndist):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    launch(2, "gpu", shape, ndist)

