"""Measure the observed-cloud producer at REAL275 resolution: CUDA producer
(catre_b200.cloud.sample_object_clouds: host->device copies of depth and masks, kernels, the one sync, host randperm,
gather) against the CPU restatement of the reference loop (oracle/cloud_oracle.py = what the data loader runs per
image today).  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from catre_b200 import cloud  # noqa: E402
from oracle import cloud_oracle as co  # noqa: E402  (CPU baseline leg only)


def scene(n_obj, seed):
    g = torch.Generator().manual_seed(seed)
    H, W = 480, 640
    K = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]], dtype=np.float32)
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 1.2 + 0.0005 * v + 0.0003 * u + 0.002 * torch.randn(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0
    masks, poses, scales = [], [], []
    for i in range(n_obj):
        cu, cv, r = 80 + 80 * i, 100 + 45 * i, 45 + 4 * i
        m = (u - cu) ** 2 + (v - cv) ** 2 <= r ** 2
        depth = torch.where(m & (depth > 0), depth - 0.4, depth)
        masks.append(m)
        z = 0.8 + 0.0005 * cv + 0.0003 * cu
        t = torch.tensor([(cu - K[0, 2]) * z / K[0, 0], (cv - K[1, 2]) * z / K[1, 1], z])
        poses.append(torch.cat((torch.eye(3), t.reshape(3, 1)), dim=1))
        scales.append(torch.tensor([0.15, 0.2, 0.18]))
    return depth.float().contiguous(), K, torch.stack(masks), torch.stack(poses).float(), torch.stack(scales).float()


def main():
    n_obj, reps = 6, 50
    depth, K, masks, poses, scales = scene(n_obj, 3)
    torch.manual_seed(1); want = co.sample_clouds(depth, K, masks, poses, scales, 1024)
    torch.manual_seed(1); got = cloud.sample_object_clouds(depth, K, masks, poses, scales, 1024)
    exact = bool(torch.equal(got.cpu(), want))
    t0 = time.perf_counter()
    for _ in range(10):
        co.sample_clouds(depth, K, masks, poses, scales, 1024)
    cpu_ms = (time.perf_counter() - t0) / 10 * 1e3
    for _ in range(5):
        cloud.sample_object_clouds(depth, K, masks, poses, scales, 1024)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        cloud.sample_object_clouds(depth, K, masks, poses, scales, 1024)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    dd, md = depth.cuda(), masks.cuda()
    t0 = time.perf_counter()
    for _ in range(reps):
        cloud.sample_object_clouds(dd, K, md, poses, scales, 1024)
    torch.cuda.synchronize()
    gpu_res_ms = (time.perf_counter() - t0) / reps * 1e3
    # device-only time of the selection kernels (CUDA events)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        cloud.select_ball_points(dd, K, md, poses, scales)
    b.record(); torch.cuda.synchronize()
    sel_ms = a.elapsed_time(b) / reps
    # several images per call (one count read-back and one sample upload per call)
    batch_ms = {}
    for nb in (4, 16):
        items = [(depth, K, masks, poses, scales)] * nb
        cloud.sample_object_clouds_batch(items, 1024)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(2, reps // nb)):
            cloud.sample_object_clouds_batch(items, 1024)
        torch.cuda.synchronize()
        batch_ms[nb] = (time.perf_counter() - t0) / (max(2, reps // nb) * nb) * 1e3
    px_bytes = n_obj * 480 * 640 * (4 + 1) * 2  # depth + mask read by the histogram and the compaction pass
    print(json.dumps({"workload": "480x640 depth, 6 objects, NUM_PCL=1024", "bit_exact_vs_cpu": exact,
                      "cpu_reference_loop_ms_per_image": round(cpu_ms, 3), "cpu_threads": torch.get_num_threads(),
                      "cuda_producer_ms_per_image_host_inputs": round(gpu_ms, 3),
                      "cuda_producer_ms_per_image_resident_inputs": round(gpu_res_ms, 3),
                      "cuda_producer_ms_per_image_host_inputs_4_images_per_call": round(batch_ms[4], 3),
                      "cuda_producer_ms_per_image_host_inputs_16_images_per_call": round(batch_ms[16], 3),
                      "select_call_ms_incl_host_radii": round(sel_ms, 4),
                      "select_algorithmic_bytes": px_bytes,
                      "speedup_vs_cpu_host_inputs": round(cpu_ms / gpu_ms, 2)}))


if __name__ == "__main__":
    main()
