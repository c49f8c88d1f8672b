#!/usr/bin/env python
"""Multi-GPU check of the tile assembly (run under torchrun, one rank per GPU):
rt_render_multi and rayito_b200::raytraceMulti() over NCCL must give rank 0 the single-GPU image
bit for bit.  Prints MULTI_CHECK_OK on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_check.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from rayito_b200 import build, capi
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    uid = torch.zeros(capi.Comm.ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    comm = capi.Comm(uid.cpu().numpy().tobytes(), rank, world, local)

    obj = build.model_path("bumpy.obj")
    host = capi.HostScene(capi.RECIPE_STAGE7_SCENE1, obj)
    scene = capi.DeviceScene(host.desc, device=local)
    spec = host.default_camera_spec()
    cam = capi.camera_from_spec(spec)
    ok = True
    for (W, H, ps, tile) in ((200, 120, 2, 0), (100, 60, 3, 16), (37, 21, 2, 64)):
        frame = torch.full((H, W, 3), -1.0, dtype=torch.float32, device=dev)
        params = capi.RtRenderParams(W, H, ps, 1, 3, tile, rank, world, 0, 0)
        stats, ms = comm.render_multi(scene, cam, params, frame.data_ptr(), 0, torch.cuda.current_stream(dev).cuda_stream)
        total = torch.tensor([float(stats.samples)], dtype=torch.float64, device=dev)
        dist.all_reduce(total)
        if rank == 0:
            whole, _ = scene.render(cam, W, H, ps, tile_size=tile)
            same = np.array_equal(frame.cpu().numpy().view(np.uint32), whole.view(np.uint32))
            print("rt_render_multi %dx%d ps%d tile %d world %d: %s, assemble %.3f ms, samples %d" % (
                W, H, ps, tile, world, "bit-identical" if same else "DIFFERENT", ms, int(total.item())), flush=True)
            ok = ok and same and int(total.item()) == W * H * ps * ps
    # the reference-facing call
    lib = capi.host()
    app = lib.rth_app_create(capi.RECIPE_STAGE7_SCENE1, obj.encode(), 0, 0)
    pixels = C.c_void_p()
    st = capi.RtRenderStats()
    W, H, ps = 160, 90, 2
    for _ in range(2):
        rc = lib.rth_app_raytrace_multi(app, spec.ctypes.data, W, H, ps, 1, 3, comm.handle, 0, C.byref(pixels), C.byref(st))
        assert rc == 0, lib.rth_last_error_string()
        if rank == 0:
            got = np.ctypeslib.as_array(C.cast(pixels, C.POINTER(C.c_float)), shape=(H, W, 3)).copy()
            whole, _ = scene.render(cam, W, H, ps)
            same = np.array_equal(got.view(np.uint32), whole.view(np.uint32))
            print("raytraceMulti %dx%d world %d: %s" % (W, H, world, "bit-identical" if same else "DIFFERENT"), flush=True)
            ok = ok and same
        else:
            assert not pixels.value
    lib.rth_app_destroy(app)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, src=0)
    scene.close()
    comm.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_CHECK_OK" if ok else "MULTI_CHECK_FAILED", flush=True)
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
