"""Dense MMA throughput of this GPU through cuBLAS (torch.matmul, 8192^3), per operand type: the denominators for
quoting tensor-pipe flop fractions (only bf16 is in MEASURED_PEAKS.json)."""
import json
import torch

dev = torch.device('cuda:0')
n = 8192
out = {}
for name, dt, tf32 in (('bf16', torch.bfloat16, False), ('fp16', torch.float16, False), ('tf32', torch.float32, True),
                       ('fp32', torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dt)
    b = torch.randn(n, n, device=dev, dtype=dt)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if name != 'fp32' else 5
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    out[name + '_tflops'] = 2 * n ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
torch.backends.cuda.matmul.allow_tf32 = False
print(json.dumps(dict(out, shape='%d^3' % n, source='torch.matmul (cuBLAS), CUDA events, 1 x B200')))
