"""the HBM-resident loader reproduces EEGDataset.__getitem__'s index arithmetic
(Retrieval/eegdatasets_leaveone.py:326-375).  Runs on CPU tensors (device='cpu'): pure host logic."""
import torch

from eeg_image_decode_b200.data import ResidentEEGData


def ref_getitem(index, train, n_cls):
    # literal restatement of the `pictures is None` branch
    if train:
        index_n_sub_train = n_cls * 10 * 4
        return (index % index_n_sub_train) // (10 * 4), (index % index_n_sub_train) // 4
    index_n_sub_test = n_cls * 1 * 80
    return index % index_n_sub_test, index % index_n_sub_test


def test_index_maps_match_reference_arithmetic():
    n_cls = 7
    for train, n in ((True, 2 * n_cls * 40), (False, n_cls)):
        eeg = torch.randn(n, 63, 250)
        labels = torch.arange(n) % n_cls
        txt = torch.randn(n_cls, 1024)
        img = torch.randn(n_cls * 10 if train else n_cls, 1024)
        ds = ResidentEEGData(eeg, labels, txt, img, train=train, n_cls=n_cls, device="cpu",
                             text=[f"t{i}" for i in range(n_cls)], img=[f"i{i}" for i in range(img.shape[0])])
        idx = torch.arange(n)
        x, lab, text, tf, im, imf = ds.batch(idx)
        for i in range(n):
            ti, ii = ref_getitem(i, train, n_cls)
            assert torch.equal(x[i], eeg[i]) and lab[i] == labels[i]
            assert torch.equal(tf[i], txt[ti]) and torch.equal(imf[i], img[ii])
            assert text[i] == f"t{ti}" and im[i] == f"i{ii}"


def test_loader_semantics():
    n_cls, n = 3, 3 * 40
    ds = ResidentEEGData(torch.randn(n, 63, 250), torch.arange(n) % n_cls, torch.randn(n_cls, 1024), torch.randn(30, 1024),
                         train=True, n_cls=n_cls, device="cpu")
    g = torch.Generator().manual_seed(0)
    batches = list(ds.loader(32, shuffle=True, drop_last=True, generator=g))
    assert len(batches) == n // 32 == len(ds.loader(32))
    seen = torch.cat([b[1] for b in batches])
    assert seen.numel() == (n // 32) * 32
    assert len(list(ds.loader(32, shuffle=False, drop_last=False))) == 4
    first = next(iter(ds.loader(8, shuffle=False)))
    assert torch.equal(first[0], ds.data[:8]) and first[2] is None and first[4] is None
