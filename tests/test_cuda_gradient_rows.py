# coding: utf-8
"""GPU parity: gradient production (`bz_gradient_row`, `GradientStack`) against the ATen operator
sequence of attack.py:776-780 (clip + clone) and :799-810 (momentum placement) run on the same GPU:
bit-exact without clipping and when the clip does not trigger; when it triggers the scale factor
`clip / norm` is an fp32 rounding of a quotient whose denominator (ATen's fp32 norm) is build
specific: 1 ulp of the scale, i.e. <= 2^-23 relative on every element."""

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
DEV = "cuda:0"

def _reference(grad, clip, mode, mom, mu, dampening):
  """ The reference's statements, verbatim semantics. """
  grad = grad.clone()
  if clip is not None:
    grad_norm = grad.norm().item()                                   # attack.py:777
    if grad_norm > clip:
      grad.mul_(clip / grad_norm)                                    # :779
  sampled = grad.clone().detach_()                                   # :780
  if mode == "worker":
    mom.mul_(mu).add_(sampled, alpha=(1. - dampening))               # :802
    return sampled, mom
  if mode == "server":
    return sampled, sampled.mul(1. - dampening).add_(mom, alpha=mu)  # :807
  return sampled, sampled

@pytest.mark.parametrize("d", [1, 31, 4099, 1_310_922])
@pytest.mark.parametrize("mode", ["update", "worker", "server"])
@pytest.mark.parametrize("clip", [None, 1e9, 2.0])
def test_gradient_row_matches_the_aten_sequence(d, mode, clip):
  import byzantinemomentum_b200 as bz
  gen = torch.Generator(device=DEV).manual_seed(d + 7)
  stack = bz.GradientStack(3, d, DEV)
  mu, damp = 0.9, 0.1
  mom_ours = torch.randn(d, device=DEV, generator=gen)
  mom_ref = mom_ours.clone()
  for i in range(3):
    grad = torch.randn(d, device=DEV, generator=gen)
    want_s, want_h = _reference(grad, clip, mode, mom_ref, mu, damp)
    kwargs = {}
    if mode == "worker":
      kwargs = dict(worker_momentum=mom_ours, mu=mu, dampening=damp)
    elif mode == "server":
      kwargs = dict(server_momentum=mom_ours, mu=mu, dampening=damp)
    got_s = stack.push(i, grad, clip=clip, **kwargs)
    got_h = mom_ours if mode == "worker" else stack.honest()[i] if mode == "server" else got_s
    triggers = clip is not None and float(grad.norm()) > clip
    if not triggers:
      assert torch.equal(got_s, want_s), (d, mode, clip, i)
      assert torch.equal(got_h, want_h), (d, mode, clip, i)
    else:
      assert torch.allclose(got_s, want_s, rtol=2.5e-7, atol=0), (d, mode, clip, i)
      assert torch.allclose(got_h, want_h, rtol=1e-6, atol=1e-7), (d, mode, clip, i)
      mom_ref.copy_(mom_ours) if mode == "worker" else None          # keep the two recursions on the same inputs
  # the rows are views of one [n, pitch] buffer: aligned, ready for the rules
  rows = stack.sampled()
  assert all(r.data_ptr() % 256 == 0 for r in rows) and rows[1].data_ptr() - rows[0].data_ptr() == stack._pitch * 4

def test_rules_read_the_stack_rows_directly():
  import byzantinemomentum_b200 as bz
  n, f, d = 11, 3, 50_021
  gen = torch.Generator(device=DEV).manual_seed(3)
  stack = bz.GradientStack(n, d, DEV)
  grads = [torch.randn(d, device=DEV, generator=gen) for _ in range(n)]
  for i, g in enumerate(grads):
    stack.push(i, g)
  for gar in ("median", "trmean", "krum", "bulyan"):
    ff = 2 if gar == "bulyan" else f
    assert torch.equal(bz.gars[gar](gradients=stack.sampled(), f=ff), bz.gars[gar](gradients=grads, f=ff))
