"""Launches each kernel family ONCE PER SHAPE at the BASELINE sizes so that `ncu --set full -k regex:...` can capture them:

    python tools/kernel_probe.py [gn|adamw|flash|flashbwd|conv|halo|sched|all] [reps]

  gn     GroupNorm(+SiLU) fwd + bwd on the config-3 level-0 / level-1 tensors (8x256x24^3, 8x512x12^3, bf16)
  adamw  fused clip + AdamW over 441 M parameters (flat buffers) + the sum-of-squares pass
  flash  flash_fwd_kernel at the config-4 attention shape (L = 32768, one 128-channel head) and the LDM shapes
  halo   conv_halo_kernel / wgrad_halo_kernel on the config-2 96^3 x 32-channel layer
  sched  ddpm_add_noise / ddpm_step / mse at a 64 MB tensor
Prints CUDA-event timings (L2 cold: every repetition uses a different buffer from a pool larger than L2)."""
import sys

import torch

sys.path.insert(0, ".")
from medical_image_generation_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda"
torch.manual_seed(0)


def timed(name, fn, nbytes=None, flops=None):
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1, reps + 1):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    extra = ""
    if nbytes:
        extra += f"  {nbytes / ms / 1e6:8.0f} GB/s (algorithmic bytes {nbytes / 1e6:.1f} MB)"
    if flops:
        extra += f"  {flops / ms / 1e9:8.0f} TFLOP/s"
    print(f"{name:58s} {ms * 1e3:9.1f} us{extra}", flush=True)


def cl(t):
    return t.contiguous(memory_format=torch.channels_last_3d)


if which in ("gn", "all"):
    for (N, C, sp, G) in ((8, 256, 24, 32), (8, 512, 12, 32), (8, 1536, 6, 32), (2, 32, 96, 16)):
        pool = [cl(torch.randn(N, C, sp, sp, sp, device=dev, dtype=torch.bfloat16)).requires_grad_(True) for _ in range(reps + 1)]
        dys = [cl(torch.randn(N, C, sp, sp, sp, device=dev, dtype=torch.bfloat16)) for _ in range(reps + 1)]
        gamma = torch.randn(C, device=dev).requires_grad_(True)
        beta = torch.randn(C, device=dev).requires_grad_(True)
        numel = pool[0].numel()
        outs = {}

        def fwd(i):
            outs[i] = ops.group_norm(pool[i], gamma, beta, G, 1e-6, silu=True)

        def bwd(i):
            outs[i].backward(dys[i])

        timed(f"groupnorm+silu fwd  {N}x{C}x{sp}^3 bf16", fwd, nbytes=2 * numel * 2)
        timed(f"groupnorm+silu bwd  {N}x{C}x{sp}^3 bf16", bwd, nbytes=3 * numel * 2)
        del pool, dys, outs

if which in ("adamw", "all"):
    from medical_image_generation_b200._lib import call
    n = 441_421_827 // 64 * 64
    master, grad = torch.randn(n, device=dev) * 0.02, torch.randn(n, device=dev) * 1e-3
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.empty(n, device=dev, dtype=torch.bfloat16)
    sumsq, partials = torch.zeros(1, device=dev), torch.zeros(2048, device=dev)
    step_dev = torch.ones(1, dtype=torch.int32, device=dev)
    st = ops._stream()
    timed("sumsq (grad-norm pass) 441 M fp32", lambda i: call("mig_sumsq", ops._ptr(grad), ops._ptr(sumsq), ops._ptr(partials), n, st),
          nbytes=4 * n)
    timed("adamw+clip+bf16 shadow 441 M params", lambda i: call(
        "mig_adamw_step", ops._ptr(master), ops._ptr(grad), ops._ptr(m), ops._ptr(v), n, 2e-5, 0.9, 0.999, 1e-8, 1e-2, 1,
        ops._ptr(sumsq), 1.0, ops._ptr(shadow), ops._ptr(step_dev), st), nbytes=30 * n)
    del master, grad, m, v, shadow

if which in ("flash", "all"):
    for (B, L, C, heads) in ((1, 32768, 128, 1), (8, 1728, 512, 1), (8, 216, 768, 1), (2, 6400, 512, 1)):
        q, k, v = (torch.randn(B, L, C, device=dev, dtype=torch.bfloat16) for _ in range(3))
        fl = 4.0 * B * L * L * C
        with torch.no_grad():
            timed(f"flash attention fwd B={B} L={L} d={C // heads}", lambda i: ops.flash_attention(q, k, v, heads, (C // heads) ** -0.5),
                  flops=fl)

if which in ("flashbwd", "all"):
    for (B, L, C, heads) in ((1, 8192, 128, 1), (8, 1728, 512, 1), (8, 216, 768, 1)):
        q, k, v = (torch.randn(B, L, C, device=dev, dtype=torch.bfloat16).requires_grad_(True) for _ in range(3))
        dO = torch.randn(B, L, C, device=dev, dtype=torch.bfloat16)
        ops.set_flash_attention(True, training=True)

        def step(i):
            ops.sdpa(q, k, v, heads, (C // heads) ** -0.5).backward(dO)

        timed(f"flash attention fwd+bwd B={B} L={L} d={C // heads}", step, flops=4.0 * B * L * L * C * 3.5)
        ops.set_flash_attention(True, training="auto")

if which in ("conv", "all"):
    # the dominant config-3 layer: Conv3d 256 -> 256, 3x3x3, 8 x 24^3 voxels -- fwd, dgrad (MN-major filter operand), wgrad
    xs = [cl(torch.randn(8, 256, 24, 24, 24, device=dev, dtype=torch.bfloat16)).requires_grad_(True) for _ in range(2)]
    w = cl(torch.randn(256, 256, 3, 3, 3, device=dev) * 0.01).requires_grad_(True)
    b = torch.zeros(256, device=dev, requires_grad=True)
    fl = 2.0 * 8 * 24 ** 3 * 256 * 256 * 27
    ys = {}

    def cfwd(i):
        ys[i % 2] = ops.conv_nd(xs[i % 2], w, b, 1, 1)

    def cbwd(i):
        ys[i % 2].backward(torch.ones_like(ys[i % 2]), retain_graph=True)

    timed("conv fwd 256->256 3^3 8x24^3 (conv_tma_kernel)", cfwd, flops=fl)
    timed("conv bwd (dgrad MN-major filter + wgrad) 256->256 8x24^3", cbwd, flops=2 * fl)
    del xs, ys

if which in ("halo", "all"):
    for (N, Cin, Cout, sp) in ((2, 32, 32, 96), (2, 32, 64, 96), (1, 32, 32, 128)):
        xs = [cl(torch.randn(N, Cin, sp, sp, sp, device=dev, dtype=torch.bfloat16)).requires_grad_(True) for _ in range(2)]
        w = cl(torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.03).requires_grad_(True)
        b = torch.zeros(Cout, device=dev, requires_grad=True)
        fl = 2.0 * N * sp ** 3 * Cin * Cout * 27
        ys = {}

        def fwd(i):
            ys[i % 2] = ops.conv_nd(xs[i % 2], w, b, 1, 1)

        def bwd(i):
            ys[i % 2].backward(torch.ones_like(ys[i % 2]), retain_graph=True)

        timed(f"conv fwd {Cin}->{Cout} 3^3 {N}x{sp}^3 (halo kernel)", fwd, flops=fl)
        timed(f"conv bwd (dgrad+wgrad) {Cin}->{Cout} {N}x{sp}^3", bwd, flops=2 * fl)
        del xs, ys

if which in ("sched", "all"):
    import medical_image_generation_b200 as mig
    from medical_image_generation_b200 import planner
    s = mig.DDPMScheduler(**planner.LDM_SCHEDULER_KWARGS)
    s.set_timesteps(1000)
    shape = (16, 1, 128, 128, 64)
    xs = [torch.randn(shape, device=dev) for _ in range(reps + 1)]
    es = [torch.randn(shape, device=dev) for _ in range(reps + 1)]
    zs = [torch.randn(shape, device=dev) for _ in range(reps + 1)]
    numel = xs[0].numel()
    t = torch.randint(0, 1000, (shape[0],), device=dev)
    timed("ddpm_add_noise fp32 16x1x128x128x64", lambda i: s.add_noise(xs[i], es[i], t), nbytes=3 * numel * 4)
    timed("ddpm_step fp32 (eps->x0, clamp, posterior, +sigma z)", lambda i: s.step(es[i], 500, xs[i], noise=zs[i]), nbytes=4 * numel * 4)
    timed("mse_loss fwd fp32", lambda i: ops.mse_loss(xs[i], es[i]), nbytes=2 * numel * 4)
