"""GPU evaluation metrics and slice preparation against the CPU oracle
(``oracle/eval_oracle.py``: torchmetrics / skimage algorithms restated in float64, scipy.ndimage called directly) and the
fixtures produced by the reference's own functions.  Metrics are fp32 on the GPU: relative tolerance 2e-4 (PSNR: 1e-3 dB);
slicing / cropping / intensity mapping is bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pair(n, h, w, seed, noise=0.05):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    base = (0.5 + 0.4 * torch.sin(5 * xx + 1) * torch.cos(4 * yy))[None] * (0.6 + 0.4 * torch.rand(n, 1, 1, generator=g))
    tgt = (base + 0.03 * torch.randn(n, h, w, generator=g)).clamp(0, 1)
    pred = (torch.nn.functional.avg_pool2d(tgt[:, None], 3, 1, 1)[:, 0] + noise * torch.randn(n, h, w, generator=g)).clamp(0, 1)
    return pred, tgt


@pytest.mark.parametrize("n,h,w", [(3, 64, 64), (2, 100, 75), (1, 512, 512), (5, 33, 47), (2, 11, 11)])
def test_image_metrics_vs_oracle(n, h, w):
    from oracle import eval_oracle as eo
    from mri_diffusion_superresolution_b200.evalmetrics import compute_mri_metrics, image_metrics

    pred, tgt = _pair(n, h, w, 100 + h)
    per, batch, sums = image_metrics(pred.cuda()[:, None], tgt.cuda()[:, None])
    per = per.cpu().numpy()
    for i in range(n):
        ref = eo.evaluator_metrics(pred[i].numpy(), tgt[i].numpy())
        assert abs(per[i, 0] - ref["PSNR"]) < 1e-3
        assert per[i, 1] == pytest.approx(ref["SSIM"], rel=2e-4, abs=2e-6)
        assert per[i, 2] == pytest.approx(ref["NMSE"], rel=2e-4)
        assert per[i, 3] == pytest.approx(ref["HFEN"], rel=2e-4)
    ps, ss, nm, hf = eo.notebook_metrics(pred[:, None].numpy(), tgt[:, None].numpy())
    got = compute_mri_metrics(pred.cuda()[:, None], tgt.cuda()[:, None])
    assert abs(got[0] - ps) < 1e-3
    assert got[1] == pytest.approx(ss, rel=2e-4, abs=2e-6)
    assert got[2] == pytest.approx(nm, rel=2e-4)
    assert got[3] == pytest.approx(hf, rel=2e-4)
    # deterministic: a second call is bit-identical (no atomics)
    per2, batch2, sums2 = image_metrics(pred.cuda()[:, None], tgt.cuda()[:, None])
    assert torch.equal(per2.cpu(), torch.from_numpy(per)) and torch.equal(batch2, batch) and torch.equal(sums2, sums)


def test_evaluator_api_and_edge_cases(golden_dir, tmp_path):
    from oracle import eval_oracle as eo
    from mri_diffusion_superresolution_b200.evalmetrics import MRIEvaluator, image_metrics

    ev = MRIEvaluator()
    pred, tgt = _pair(4, 96, 80, 7)
    res = ev.evaluate_pairs(pred.cuda(), tgt.cuda())
    refs = [eo.evaluator_metrics(pred[i].numpy(), tgt[i].numpy()) for i in range(4)]
    for k in ("PSNR", "SSIM", "HFEN", "NMSE"):
        assert res[k] == pytest.approx(np.mean([r[k] for r in refs]), rel=2e-4)     # mean over pairs (not / 13, eval.py:91)
    assert ev.compute_hfen(pred[0], tgt[0]) == pytest.approx(refs[0]["HFEN"], rel=2e-4)
    assert ev.compute_hfen(pred[0].numpy(), tgt[0].numpy(), sigma=1.0) == pytest.approx(eo.hfen(pred[0].numpy(), tgt[0].numpy(), 1.0), rel=2e-4)
    assert ev.compute_nmse(pred[1][None, None], tgt[1][None, None]) == pytest.approx(refs[1]["NMSE"], rel=2e-4)
    assert float(ev.ssim(pred.cuda()[:, None], tgt.cuda()[:, None])) == pytest.approx(np.mean([r["SSIM"] for r in refs]), rel=2e-4)
    # the reference's compute_nmse on its own fixture
    z = np.load(os.path.join(golden_dir, "eval_metrics.npz"))
    per, _, _ = image_metrics(torch.from_numpy(z["pred"]).cuda(), torch.from_numpy(z["target"]).cuda())
    np.testing.assert_allclose(per[:, 2].cpu().numpy(), z["nmse"], rtol=2e-4)
    # identical images: SSIM 1, NMSE 0, HFEN 0, PSNR inf
    per, _, _ = image_metrics(tgt.cuda(), tgt.cuda())
    per = per.cpu().numpy()
    assert np.all(np.isinf(per[:, 0])) and np.allclose(per[:, 1], 1.0, atol=1e-6) and np.all(per[:, 2:] == 0)
    with pytest.raises(ValueError):
        ev.compute_hfen(pred[0], tgt[0], sigma=3.0)             # radius 12 > 6
    with pytest.raises(ValueError):
        image_metrics(pred.cuda()[:, :8, :8].contiguous(), tgt.cuda()[:, :8, :8].contiguous())
    with pytest.raises(RuntimeError):
        image_metrics(pred, tgt)
    # evaluate_folders: png pairs on disk
    import cv2
    gd, td = tmp_path / "gen", tmp_path / "gt"
    gd.mkdir(), td.mkdir()
    imgs = []
    for i in range(3):
        a = (pred[i].numpy() * 255).astype(np.uint8)
        b = (tgt[i].numpy() * 255).astype(np.uint8)
        cv2.imwrite(str(gd / f"s{i:02d}.png"), a)
        cv2.imwrite(str(td / f"s{i:02d}.png"), b)
        imgs.append((a.astype(np.float32) / 255.0, b.astype(np.float32) / 255.0))
    res = ev.evaluate_folders(str(gd), str(td))
    for k in ("PSNR", "SSIM", "HFEN", "NMSE"):
        assert res[k] == pytest.approx(np.mean([eo.evaluator_metrics(a, b)[k] for a, b in imgs]), rel=2e-4)


def test_volume_slicing_bit_exact(golden_dir):
    from oracle import eval_oracle as eo
    from mri_diffusion_superresolution_b200.slices import pad_or_center_crop, volume_to_slices

    g = np.random.default_rng(3)
    for shape, clip in (((512, 512, 128), (0.0, 2000.0)), ((300, 470, 37), (0.0, 900.0)), ((600, 530, 5), (10.0, 700.0)),
                        ((700, 128, 33), (0.0, 1000.0))):
        v = (g.random(shape, dtype=np.float32) * 1.3 * clip[1] - 50).astype(np.float32)
        ref = eo.volume_to_slices(v, clip[0], clip[1])
        out = volume_to_slices(torch.from_numpy(v).cuda(), clip[0], clip[1])
        assert tuple(out.shape) == ref.shape
        np.testing.assert_array_equal(out.cpu().numpy(), ref)                                  # bit-exact
    z = np.load(os.path.join(golden_dir, "eval_metrics.npz"))
    for tag in ("small", "big", "mixed", "exact"):
        res = pad_or_center_crop(torch.from_numpy(z[f"crop_in_{tag}"].astype(np.float32) / 64.0).cuda()).cpu().numpy()
        assert list(res.shape) == z[f"crop_out_shape_{tag}"].tolist()
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(res).tobytes()).digest(), dtype=np.uint8)
        np.testing.assert_array_equal(digest, z[f"crop_out_sha256_{tag}"])                      # the reference's own output
    with pytest.raises(ValueError):
        volume_to_slices(torch.zeros(4, 4, device="cuda"), 0, 1)
    with pytest.raises(ValueError):
        volume_to_slices(torch.zeros(4, 4, 4, device="cuda"), 1.0, 1.0)


def test_mnist_toy_pieces_vs_reference_golden(golden_dir):
    """The runnable cells of the MNIST notebook (forward_pass, SinusoidalPositionEmbeddings) on the GPU vs the notebook's
    own output.  forward_pass is bit-exact; the embedding is fp32 sin / cos of arguments up to 999 rad, where one ulp of
    the frequency moves the result by 6e-5: tolerance 2e-4 absolute."""
    from mri_diffusion_superresolution_b200 import mnist

    z = np.load(os.path.join(golden_dir, "mnist_toy.npz"))
    x0, noise = torch.from_numpy(z["x0"]).cuda(), torch.from_numpy(z["noise"]).cuda()
    np.testing.assert_array_equal(mnist.forward_pass(x0, 321, noise).cpu().numpy(), z["fwd_scalar"])
    np.testing.assert_array_equal(mnist.forward_pass(x0, torch.from_numpy(z["t_vec"]), noise).cpu().numpy(), z["fwd_vec"])
    t = torch.from_numpy(z["t_emb"]).cuda()
    e32 = mnist.SinusoidalPositionEmbeddings(32)(t)
    assert tuple(e32.shape) == (5, 32) and e32.dtype == torch.float32
    np.testing.assert_allclose(e32.cpu().numpy(), z["emb32"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(mnist.SinusoidalPositionEmbeddings(64)(t.float()).cpu().numpy(), z["emb64"], rtol=0, atol=2e-4)
    np.testing.assert_array_equal(e32[0].cpu().numpy(), np.r_[np.zeros(16), np.ones(16)].astype(np.float32))   # [sin | cos] at t = 0
    with pytest.raises(RuntimeError):
        mnist.forward_pass(x0.cpu(), 3, noise.cpu())
