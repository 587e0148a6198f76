"""CPU, world_size 2 over gloo: the view-sharded step — sharding, gradient bucket,
allreduce — with the oracle-backed stand-in as the renderer.  Summed gradients must equal a
single-process pass over all views."""
import os
import socket

import torch
import torch.multiprocessing as mp

N_VIEWS = 4
P = 300
W, H = 40, 24


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    from workload import synthetic
    from bloomscene_b200.multiview import GaussianParams
    from bloomscene_b200.rasterizer import bind
    from oracle_backend import OracleBackend

    torch.set_num_threads(1)
    api = bind(OracleBackend(threads=1))
    scene = synthetic.make_scene(P, "object", "sh1", -2.8, seed=5)
    cams = [synthetic.orbit_camera(W, H, 2 * 3.14159265 * k / N_VIEWS) for k in range(N_VIEWS)]
    Wc, Wd = synthetic.loss_weights(W, H, seed=2)
    loss_fn = lambda color, depth, vi: (color * Wc).sum() * (1.0 + 0.1 * vi) + (depth * Wd).sum()
    return api, GaussianParams(scene), cams, loss_fn


def _worker(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from bloomscene_b200.multiview import view_sharded_step

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api, params, cams, loss_fn = _setup()
    res = view_sharded_step(params, cams, torch.zeros(3), api.GaussianRasterizer, loss_fn, rank=rank, world=world)
    # sharded host staging: every rank uploads / downloads only its slice, the rest travels rank to rank
    from bloomscene_b200.multiview import download_grads, upload_params

    host_grads = torch.zeros_like(params.grad_bucket)
    down = download_grads(params, host_grads, rank, world)
    staged = {}
    for trim in (0, 1):  # numel divisible by the world size (all-gather) and not (per-slice broadcasts)
        n = params.flat.numel() - trim
        host_params = torch.arange(params.flat.numel(), dtype=torch.float32) * 0.5 + 1.0
        saved = params.flat.detach().clone()
        if trim:
            class _Trim:  # a parameter set whose flat buffer is one element shorter
                flat = params.flat.detach()[:n]
            up = upload_params(_Trim, host_params[:n], rank, world)
        else:
            up = upload_params(params, host_params, rank, world)
        staged[trim] = (up, torch.equal(params.flat.detach()[:n], host_params[:n]))
        with torch.no_grad():
            params.flat.copy_(saved)
    torch.save({"bucket": params.grad_bucket.clone(), "loss": res["loss"], "views": res["views"],
                "host_grads": host_grads, "down": down, "staged": staged},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_step_matches_single_process(tmp_path):
    from bloomscene_b200.multiview import shard_views, view_sharded_step

    assert shard_views(7, 1, 3) == [1, 4]
    assert sorted(shard_views(5, 0, 2) + shard_views(5, 1, 2)) == list(range(5))

    api, params, cams, loss_fn = _setup()
    single = view_sharded_step(params, cams, torch.zeros(3), api.GaussianRasterizer, loss_fn)
    ref_bucket = params.grad_bucket.clone()
    assert ref_bucket.abs().sum() > 0
    # parameter grads are views of the bucket (one allreduce covers all of them)
    assert params.tensors["means3D"].grad.data_ptr() == params.grad_bucket.data_ptr()

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert r0["views"] == [0, 2] and r1["views"] == [1, 3]
    assert torch.equal(r0["bucket"], r1["bucket"])  # allreduce leaves identical buckets on every rank
    rel = (r0["bucket"].double() - ref_bucket.double()).norm() / ref_bucket.double().norm()
    assert rel <= 1e-5
    assert abs(r0["loss"].item() - single["loss"].item()) <= 1e-4 * abs(single["loss"].item())
    # host staging: the two ranks' downloads tile the reduced bucket, uploads reproduce the host copy everywhere
    n = ref_bucket.numel()
    assert r0["down"] + r1["down"] == 4 * n
    assert torch.equal(r0["host_grads"] + r1["host_grads"], r0["bucket"])
    for r in (r0, r1):
        for trim in (0, 1):
            up, same = r["staged"][trim]
            assert same and up < 4 * n
