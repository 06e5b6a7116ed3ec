"""Surrogate loss heads of the train step (SURVEY.md §8f-3 / §8d cfg 4): architecture and determinism checks on the CPU."""
import torch

from stylegan_directions_face_reenactment_b200.loss_heads import ArcFaceIRSE50, LPIPSAlex, SurrogateLossHeads


def test_heads_are_seeded_and_differentiable():
    a, b = SurrogateLossHeads(amp=False), SurrogateLossHeads(amp=False)
    for (n1, p1), (n2, p2) in zip(a.state_dict().items(), b.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2), n1                     # identical replicas on every rank
    assert not any(p.requires_grad for p in a.parameters())            # frozen, like the reference's .eval() heads
    g = torch.Generator().manual_seed(0)
    shifted = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).requires_grad_(True)
    source = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    target = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    loss, parts = a(shifted, source, target)
    loss.backward()
    assert torch.isfinite(loss) and set(parts) == {'loss_identity', 'loss_perceptual', 'loss_shape'}
    assert shifted.grad is not None and torch.isfinite(shifted.grad).all() and shifted.grad.abs().max() > 0
    # identical images: identity and perceptual terms vanish
    same, parts = a(source, source, source)
    assert float(parts['loss_identity']) < 1e-5 and float(parts['loss_perceptual']) < 1e-8 and float(parts['loss_shape']) < 1e-6


def test_architectures_match_the_reference_layouts():
    net = ArcFaceIRSE50()
    # Backbone(112, 50, 'ir_se') of libs/criteria/model_irse.py: 3 + 4 + 14 + 3 units, 43.8 M parameters, 512-d unit-norm output
    assert len(net.body) == 24 and sum(p.numel() for p in net.parameters()) == 43797696
    net.eval()
    with torch.no_grad():
        e = net(ArcFaceIRSE50.crop(torch.randn(2, 3, 256, 256)))
    assert e.shape == (2, 512) and torch.allclose(e.norm(dim=1), torch.ones(2), atol=1e-5)
    lp = LPIPSAlex().eval()
    with torch.no_grad():
        taps = lp.taps(torch.randn(1, 3, 256, 256))
    assert [t.shape[1] for t in taps] == [64, 192, 384, 256, 256]       # lpips/networks.py:91-98
