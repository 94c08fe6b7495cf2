"""GPU: the torch.library registration of the hot-path operators (namespace medimgen_b200): dispatcher-visible schema,
CUDA implementation = the C-ABI call, fake implementation for shape propagation, autograd formula of C-ABI kernels."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import BF16_TOL, FP32_TOL, bf16_round, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last_3d if t.ndim == 5 else torch.channels_last)


def test_ops_are_registered_with_schemas():
    import medical_image_generation_b200.custom_ops  # noqa: F401
    ns = torch.ops.medimgen_b200
    for name in ("conv_nd", "group_norm", "sdpa", "ddpm_add_noise", "ddpm_step", "mse_loss"):
        op = getattr(ns, name)
        assert "Tensor" in str(op.default._schema), name
    assert "[] stride" in str(ns.conv_nd.default._schema)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_conv_nd_custom_op_matches_reference_with_grads(dtype, tol):
    import medical_image_generation_b200.custom_ops  # noqa: F401
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 8, 8, 8, generator=g)
    w = torch.randn(128, 64, 3, 3, 3, generator=g) / math.sqrt(64 * 27)
    b = torch.randn(128, generator=g) * 0.1
    if dtype == torch.bfloat16:
        x, w = bf16_round(x), bf16_round(w)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y_ref = F.conv3d(xr, wr, br, stride=1, padding=1)
    probe = torch.randn(y_ref.shape, generator=g)
    (y_ref * probe).sum().backward()
    xd = _cl(x.to(DEV).to(dtype)).requires_grad_(True)
    wd = _cl(w.to(DEV)).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    y = torch.ops.medimgen_b200.conv_nd(xd, wd, bd, None, None, [1, 1, 1], [1, 1, 1])
    (y.float() * probe.to(DEV)).sum().backward()
    assert rel_err(y, y_ref) < tol and rel_err(xd.grad, xr.grad) < tol
    assert rel_err(wd.grad, wr.grad) < tol and rel_err(bd.grad, br.grad) < tol


def test_group_norm_and_sdpa_custom_ops():
    import medical_image_generation_b200.custom_ops  # noqa: F401
    g = torch.Generator().manual_seed(2)
    x = bf16_round(torch.randn(2, 64, 6, 6, 6, generator=g))
    gamma, beta = 1 + 0.2 * torch.randn(64, generator=g), 0.2 * torch.randn(64, generator=g)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    y_ref = F.silu(F.group_norm(xr, 16, gr, br, 1e-6))
    probe = bf16_round(torch.randn(y_ref.shape, generator=g))
    (y_ref * probe).sum().backward()
    xd = _cl(x.to(DEV).bfloat16()).requires_grad_(True)
    gd, bd = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y, mean, rstd = torch.ops.medimgen_b200.group_norm(xd, gd, bd, 16, 1e-6, True)
    assert mean.shape == (2, 16) and rstd.shape == (2, 16)
    y.backward(_cl(probe.to(DEV).bfloat16()))
    assert rel_err(y, y_ref) < BF16_TOL and rel_err(xd.grad, xr.grad) < BF16_TOL
    assert rel_err(gd.grad, gr.grad) < BF16_TOL and rel_err(bd.grad, br.grad) < BF16_TOL
    # fused attention, forward + backward through the dispatcher
    q, k, v = (bf16_round(torch.randn(2, 200, 128, generator=g)) for _ in range(3))
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    want = F.scaled_dot_product_attention(qr[:, None], kr[:, None], vr[:, None])[:, 0]
    pr = bf16_round(torch.randn(want.shape, generator=g))
    (want * pr).sum().backward()
    qd, kd, vd = (t.to(DEV).bfloat16().requires_grad_(True) for t in (q, k, v))
    out, lse = torch.ops.medimgen_b200.sdpa(qd, kd, vd, 1, 1 / math.sqrt(128))
    assert lse.shape == (2, 200)
    (out.float() * pr.to(DEV)).sum().backward()
    assert rel_err(out, want) < BF16_TOL
    assert rel_err(qd.grad, qr.grad) < BF16_TOL and rel_err(kd.grad, kr.grad) < BF16_TOL and rel_err(vd.grad, vr.grad) < BF16_TOL


def test_scheduler_and_loss_custom_ops_and_opcheck():
    import medical_image_generation_b200 as mig
    import medical_image_generation_b200.custom_ops  # noqa: F401
    from oracle.ddpm_oracle import OracleDDPMScheduler
    kw = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0205)
    s, o = mig.DDPMScheduler(**kw), OracleDDPMScheduler(**kw)
    g = torch.Generator().manual_seed(3)
    x0, nz, z = (torch.randn(2, 3, 6, 6, 6, generator=g) for _ in range(3))
    t = torch.tensor([10, 700])
    noisy = torch.ops.medimgen_b200.ddpm_add_noise(x0.to(DEV), nz.to(DEV), t.to(DEV), s.alphas_cumprod.to(DEV), False)
    assert rel_err(noisy, o.add_noise(x0, nz, t)) < 1e-6
    kc = s.step_coefficients(700)
    prev, x0h = torch.ops.medimgen_b200.ddpm_step(nz.to(DEV), noisy, z.to(DEV), kc["sqrt_acp"], kc["sqrt_one_minus_acp"],
                                                  kc["c0"], kc["ct"], kc["sigma"], 0, True)
    wprev, wx0 = o.step(nz, 700, noisy.cpu(), noise=z)
    assert rel_err(prev, wprev) < 1e-5 and rel_err(x0h, wx0) < 1e-5
    a = torch.randn(4, 5, 6, device=DEV, requires_grad=True)
    b = torch.randn(4, 5, 6, device=DEV)
    loss = torch.ops.medimgen_b200.mse_loss(a, b, False)
    loss.backward()
    assert rel_err(loss, F.mse_loss(a.detach(), b)) < 1e-6 and rel_err(a.grad, 2 * (a.detach() - b) / a.numel()) < 1e-5
    # schema / fake-tensor / autograd-registration checks of torch.library
    xd = _cl(torch.randn(1, 16, 4, 4, 4, device=DEV)).requires_grad_(True)
    wd = _cl(torch.randn(16, 16, 3, 3, 3, device=DEV) * 0.1).requires_grad_(True)
    torch.library.opcheck(torch.ops.medimgen_b200.conv_nd.default, (xd, wd, None, None, None, [1, 1, 1], [1, 1, 1]),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.medimgen_b200.mse_loss.default, (a.detach().requires_grad_(True), b, False),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
