"""The client's input path (fedcola_b200/client/fedavgclient.py): batch order and RNG consumption equal to the
reference's DataLoader iteration (fedavgclient.py:44-53,79), per-batch double-buffered host->device feeding, and no
frozen train-time transforms."""
import pytest
import torch

from fedcola_b200.client.fedavgclient import _ClientData, epoch_batches


class _Items(torch.utils.data.Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return torch.full((3,), float(i)), i


@pytest.mark.parametrize("shuffle", [True, False])
def test_epoch_batches_draw_like_the_reference_dataloader(shuffle):
    """Reference: `for inputs, targets in DataLoader(training_set, batch_size=B, shuffle=not no_shuffle)`.
    Same permutations for two consecutive epochs AND the same global RNG state afterwards (DropPath masks and every
    later draw depend on it)."""
    n, B = 37, 8
    ds = _Items(n)
    torch.manual_seed(1)
    ref_loader = torch.utils.data.DataLoader(dataset=ds, batch_size=B, shuffle=shuffle)
    want = [[y.tolist() for _, y in ref_loader] for _ in range(2)]
    state_ref = torch.get_rng_state()
    torch.manual_seed(1)
    idx_loader = torch.utils.data.DataLoader(dataset=range(n), batch_size=B, shuffle=shuffle)
    got = [epoch_batches(idx_loader) for _ in range(2)]
    assert got == want
    assert torch.equal(torch.get_rng_state(), state_ref)
    if shuffle:
        assert want[0] != want[1] and sorted(sum(want[0], [])) == list(range(n))


class _Tensors(torch.utils.data.Dataset):
    def __init__(self, n):
        g = torch.Generator().manual_seed(0)
        self.x = torch.randn(n, 3, 8, 8, generator=g).double()       # wrong dtype on purpose: must be cast
        self.y = torch.randint(0, 10, (n,), generator=g).int()

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class _Jitter(torch.utils.data.Dataset):
    """A stochastic train-time transform: every __getitem__ draws fresh noise."""

    def __len__(self):
        return 12

    def __getitem__(self, i):
        return torch.full((3, 4, 4), float(i)) + torch.rand(3, 4, 4), i


@pytest.mark.gpu
@pytest.mark.parametrize("resident", ["host", "device"])
def test_feed_delivers_the_indexed_rows(resident, cuda):
    ds = _Tensors(50)
    data = _ClientData(ds, "img", resident, cuda)
    batches = [[5, 6, 7, 8], [40, 2, 3, 4, 49, 0, 17], [9], list(range(20, 36)), [1, 48]]
    seen = []
    for idx, (a, b) in zip(batches, data.feed(batches)):
        assert a.dtype == torch.float32 and b.dtype == torch.int64 and a.is_cuda and a.is_contiguous()
        seen.append((a.clone(), b.clone()))          # the slot is recycled two batches later
    torch.cuda.synchronize()
    for idx, (a, b) in zip(batches, seen):
        assert torch.equal(a.cpu(), ds.x[idx].float()) and torch.equal(b.cpu(), ds.y[idx].long())


@pytest.mark.gpu
def test_generic_datasets_are_rematerialised_every_epoch(cuda):
    """caching __getitem__ results froze RandomCrop/ColorJitter-style transforms to one draw per sample."""
    data = _ClientData(_Jitter(), "img", "host", cuda)
    assert data.cols is None
    batches = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]]
    epochs = []
    for _ in range(2):
        epochs.append(torch.cat([a.clone() for a, _ in data.feed(batches)]).cpu())
    assert not torch.equal(epochs[0], epochs[1])                       # fresh draws
    assert torch.equal(epochs[0].floor(), epochs[1].floor())           # same samples, same order
    cached = _ClientData(_Jitter(), "img", "host", cuda, cache_items=True)   # explicit opt-in keeps one draw
    e = [torch.cat([a.clone() for a, _ in cached.feed(batches)]).cpu() for _ in range(2)]
    assert torch.equal(e[0], e[1])
