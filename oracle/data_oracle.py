"""TEST INFRASTRUCTURE ONLY: the reference's spiking-dataset item / batch construction restated in numpy
(sparch/dataloaders/spiking_datasets.py:66-86).  Pinned in tests/test_host_contract.py against the reference's own
``SpikingDataset.__getitem__`` / ``generateBatch`` (build container); the checker of sparch_b200.data on the GPU."""
import numpy as np


def time_bins(nb_steps=100, max_time=1.4):
    return np.linspace(0, max_time, num=nb_steps)                        # spiking_datasets.py:54


def example_to_dense(times, units, nb_steps=100, nb_units=700, max_time=1.4):
    """spiking_datasets.py:68-78: digitize, sparse tensor of ones at (bin, unit), to_dense() -- duplicate indices SUM,
    the dense tensor holds spike counts.  Raises IndexError where the sparse constructor would (bin == nb_steps)."""
    t = np.digitize(times, time_bins(nb_steps, max_time))                # :68
    u = np.asarray(units, np.int64)
    if len(t) and (t.max() >= nb_steps or u.min() < 0 or u.max() >= nb_units):
        raise IndexError("event outside the (nb_steps, nb_units) grid")
    x = np.zeros((nb_steps, nb_units), np.float32)
    np.add.at(x, (t, u), 1.0)
    return x


def batch_to_dense(times_list, units_list, labels, **kw):
    """generateBatch (spiking_datasets.py:80-86): pad_sequence over equal-length items = stack."""
    xs = np.stack([example_to_dense(t, u, **kw) for t, u in zip(times_list, units_list)])
    xlens = np.full(len(times_list), xs.shape[1], np.int64)
    return xs, xlens, np.asarray(labels, np.int64)


def synthetic_events(n_examples, seed=0, rate=8000, nb_units=700, max_time=1.4):
    """SHD-shaped synthetic event lists: ~rate events per example, times as the files store them (float16 seconds,
    strictly inside [0, max_time)), units uint16."""
    rng = np.random.default_rng(seed)
    T, U = [], []
    for _ in range(n_examples):
        n = int(rng.integers(rate // 2, rate * 3 // 2))
        t = np.sort(rng.uniform(0, max_time * 0.98, n)).astype(np.float16)
        T.append(t)
        U.append(rng.integers(0, nb_units, n).astype(np.uint16))
    return T, U, rng.integers(0, 20, n_examples)
