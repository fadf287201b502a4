"""Dense TF32 tensor throughput of this GPU as cuBLAS delivers it (SURVEY 8d asks for a measured TF32 peak beside the
bf16 one in MEASURED_PEAKS.json): torch.matmul of fp32 8192^3 with TF32 allowed, best of 10 (burst) and back to back
for 3 s (sustained); the same for bf16 as a cross-check of MEASURED_PEAKS.json. One JSON line."""
import json
import time

import torch


def measure(dtype, allow_tf32):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    n = 8192
    a = torch.randn((n, n), device='cuda', dtype=dtype)
    b = torch.randn((n, n), device='cuda', dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    flop = 2.*n**3
    best = 0.
    for _ in range(10):
        (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop/(e0.elapsed_time(e1)*1e-3)/1e12)
    (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    count = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 3.:
        for _ in range(20):
            torch.matmul(a, b)
        count += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return (best, flop*count/(e0.elapsed_time(e1)*1e-3)/1e12)


def main():
    (tf32_burst, tf32_sustained) = measure(torch.float32, True)
    (bf16_burst, bf16_sustained) = measure(torch.bfloat16, False)
    print(json.dumps({'gpu': torch.cuda.get_device_name(0), 'tf32_tflops': tf32_burst, 'tf32_tflops_sustained': tf32_sustained,
                      'bf16_tflops': bf16_burst, 'bf16_tflops_sustained': bf16_sustained,
                      'how': 'torch.matmul 8192^3 (2 N^3 flop), fp32 inputs with torch.backends.cuda.matmul.allow_tf32 = True '
                             '(cuBLAS TF32 kernels) and bf16; best of 10 (burst), back to back for 3 s (sustained); CUDA events'}))


if __name__ == '__main__':
    main()
