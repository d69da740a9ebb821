"""Deterministic synthetic molecules in the reference's input layout:
`atoms` int32 (mb, N) atomic numbers with 0 = padding, `adj` float32 (mb, 4, N, N)
symmetric 0/1 per bond type (single, double, triple, aromatic), no self loops --
what `construct_discrete_edge_matrix` + `concat_mols` hand the model
(my_utils/preprocessors/ggnn_preprocessor.py:77, train_binary.py:33,553).

Molecules are random chain-biased spanning trees plus floor(n/6) ring closures,
degree capped at 4.  Vectorised over molecules (the loop is over atom index)."""
import numpy as np

ATOM_IDS = np.array([6, 7, 8, 9, 15, 16, 17, 35], dtype=np.int32)
ATOM_P = np.array([0.70, 0.10, 0.12, 0.02, 0.01, 0.02, 0.02, 0.01])
BOND_P = np.array([0.70, 0.12, 0.02, 0.16])
DATA_SEED = 2018      # setting.py:28 GLOBAL_SEED
PARAM_SEED = 777      # train_binary.py:381 --seed default


def random_molecules(rng, n_mol, n_max, n_min=None, pad_to=None):
    n_min = max(2, n_max // 2) if n_min is None else n_min
    N = n_max if pad_to is None else pad_to
    n = rng.integers(n_min, n_max + 1, size=n_mol)
    idx = np.arange(N)[None, :]
    live = idx < n[:, None]
    atoms = np.where(live, rng.choice(ATOM_IDS, size=(n_mol, N), p=ATOM_P), 0).astype(np.int32)
    adj = np.zeros((n_mol, 4, N, N), dtype=np.float32)
    deg = np.zeros((n_mol, N), dtype=np.int32)
    mols = np.arange(n_mol)

    def bond(sel, a, b, t):
        adj[sel, t, a, b] = 1.0
        adj[sel, t, b, a] = 1.0
        np.add.at(deg, (sel, a), 1)
        np.add.at(deg, (sel, b), 1)

    for i in range(1, n_max):
        act = i < n
        parent = np.where(rng.random(n_mol) < 0.7, i - 1, rng.integers(0, i, size=n_mol))
        full = deg[mols, parent] >= 4
        first_ok = np.argmax(deg[:, :i] < 4, axis=1)
        parent = np.where(full, first_ok, parent)
        t = rng.choice(4, size=n_mol, p=BOND_P)
        sel = mols[act]
        bond(sel, np.full(sel.shape, i), parent[act], t[act])
    for _ in range(n_max // 6):
        a = (rng.random(n_mol) * n).astype(np.int64)
        b = (rng.random(n_mol) * n).astype(np.int64)
        ok = (a != b) & (deg[mols, a] < 4) & (deg[mols, b] < 4) & (adj[mols, :, a, b].sum(axis=1) == 0)
        t = rng.choice(4, size=n_mol, p=BOND_P)
        sel = mols[ok]
        bond(sel, a[ok], b[ok], t[ok])
    return atoms, adj


def random_pairs(seed, n_pairs, n_max=64, n_classes=1, pad_to=None, n_max2=None):
    """(atoms_1, adj_1, atoms_2, adj_2, labels int32 (n_pairs, n_classes))."""
    rng = np.random.default_rng(seed)
    a1, A1 = random_molecules(rng, n_pairs, n_max, pad_to=pad_to)
    a2, A2 = random_molecules(rng, n_pairs, n_max2 or n_max, pad_to=pad_to)
    if n_classes == 1:
        y = (rng.random((n_pairs, 1)) < 0.33).astype(np.int32)      # pos-rate of RECORD.txt:38
    else:
        y = np.zeros((n_pairs, n_classes), dtype=np.int32)
        y[np.arange(n_pairs), rng.integers(0, n_classes, size=n_pairs)] = 1
    return a1, A1, a2, A2, y
