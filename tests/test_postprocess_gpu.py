"""Device-side connected components / quads (csrc/postprocess.cu) vs scipy.ndimage.label (8-connectivity) and the
reference's extract_cc_quads recipe (cv2.findContours RETR_EXTERNAL + minAreaRect + boxPoints, postprocess.py:27-35)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _blobs(g, H, W, n_rect=25):
    m = torch.zeros(H, W)
    for _ in range(n_rect):
        h, w = int(torch.randint(2, min(30, H // 2), (1,), generator=g)), int(torch.randint(2, min(60, W // 2), (1,), generator=g))
        y, x = int(torch.randint(0, H - h, (1,), generator=g)), int(torch.randint(0, W - w, (1,), generator=g))
        m[y : y + h, x : x + w] = 1.0
    noise = torch.rand(H, W, generator=g)
    m[noise < 0.02] = 1.0   # isolated pixels and diagonal touches (8-connectivity matters)
    m[noise > 0.985] = 0.0  # holes
    return m


@pytest.mark.parametrize("N,H,W", [(3, 200, 152), (1, 37, 51), (2, 600, 800)])
def test_labels_match_scipy_8_connectivity(N, H, W):
    import scipy.ndimage as ndi

    from ocrs_models_b200.postprocess import connected_components

    g = torch.Generator().manual_seed(H)
    masks = torch.stack([_blobs(g, H, W) for _ in range(N)])
    labels, ncomp = connected_components(masks.cuda())
    labels, ncomp = labels.cpu().numpy(), ncomp.cpu().numpy()
    for n in range(N):
        ref, k = ndi.label(masks[n].numpy() > 0.5, structure=np.ones((3, 3)))
        assert ncomp[n] == k
        assert ((labels[n] > 0) == (ref > 0)).all()
        # same partition: the map ref-label -> our-label is a bijection, and ours = 1 + min pixel index of the component
        pairs = np.unique(np.stack([ref[ref > 0], labels[n][ref > 0]]), axis=1)
        assert pairs.shape[1] == k and len(np.unique(pairs[0])) == k and len(np.unique(pairs[1])) == k
        flat = labels[n].ravel()
        first = {int(l): int(np.flatnonzero(flat == l)[0]) for l in np.unique(flat) if l > 0}
        assert all(l == i + 1 for l, i in first.items())


def _canon(quads):
    """Rectangles as sorted corner lists, sorted by centre (cv2 may start at another corner / return another order)."""
    q = np.asarray(quads, dtype=np.float64).reshape(-1, 4, 2)
    q = np.stack([c[np.lexsort((c[:, 1], c[:, 0]))] for c in q]) if len(q) else q
    return q[np.lexsort((q[:, :, 1].mean(1), q[:, :, 0].mean(1)))] if len(q) else q


def test_quads_match_the_reference_recipe():
    import cv2

    from ocrs_models_b200.postprocess import batch_cc_quads, extract_cc_quads

    g = torch.Generator().manual_seed(5)
    masks = torch.stack([_blobs(g, 300, 400, n_rect=12) for _ in range(2)])
    ours = batch_cc_quads(masks.cuda())
    for n in range(2):
        contours, _ = cv2.findContours(masks[n].to(torch.uint8).numpy(), mode=cv2.RETR_EXTERNAL, method=cv2.CHAIN_APPROX_SIMPLE)
        ref = np.array([cv2.boxPoints(cv2.minAreaRect(c[:, 0])) for c in contours])  # postprocess.py:31-35
        a, b = _canon(ours[n].numpy()), _canon(ref)
        assert a.shape == b.shape
        # areas and centres agree; corners agree where the minimum-area rectangle is unique
        np.testing.assert_allclose(a.mean(1), b.mean(1), atol=0.51)

        def area(q):  # shoelace over the corners in the order cv2.boxPoints returned them
            x, y = q[:, :, 0], q[:, :, 1]
            return 0.5 * np.abs((x * np.roll(y, -1, 1) - np.roll(x, -1, 1) * y).sum(1))

        oa = area(np.asarray(ours[n].numpy(), dtype=np.float64))
        ra = area(np.asarray(ref, dtype=np.float64))
        np.testing.assert_allclose(np.sort(oa), np.sort(ra), rtol=1e-4, atol=1e-3)  # the minimum area itself is unique
    single = extract_cc_quads(masks[0][None].cuda())
    assert single.shape == ours[0].shape
