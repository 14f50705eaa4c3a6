"""TEST INFRASTRUCTURE ONLY.  An independent solution of the rigid-molecule constraint problem that
oracle/constv_oracle_settle_impl.h solves analytically: iterative SHAKE (Ryckaert, Ciccotti & Berendsen,
J. Comput. Phys. 23, 327 (1977)) in float64, converged to rounding.  SETTLE and converged SHAKE define the
same constrained positions (new = unconstrained + sum_ij lambda_ij * (old bond vector ij) / m_i with the
three distances restored), so agreement between the two pins the oracle's algebra without OpenMM."""
import numpy as np


def shake_molecules(old, new, masses, dist_ab, dist_bc, tol=1e-15, max_iter=2000):
    """old, new: (n, 3, 3) positions (apex, b, c); masses (n, 3); returns constrained new positions."""
    old = np.asarray(old, np.float64)
    r = np.array(new, np.float64)
    m = np.asarray(masses, np.float64)
    pairs = ((0, 1, dist_ab), (0, 2, dist_ab), (1, 2, dist_bc))
    for _ in range(max_iter):
        worst = 0.0
        for i, j, d in pairs:
            d = np.asarray(d, np.float64)
            s = r[:, i] - r[:, j]
            diff = d * d - np.einsum("nk,nk->n", s, s)
            worst = max(worst, float(np.max(np.abs(diff) / (d * d))))
            ref = old[:, i] - old[:, j]
            g = diff / (2.0 * (1.0 / m[:, i] + 1.0 / m[:, j]) * np.einsum("nk,nk->n", s, ref))
            r[:, i] += (g / m[:, i])[:, None] * ref
            r[:, j] -= (g / m[:, j])[:, None] * ref
        if worst < tol:
            break
    return r


def find_settle_clusters(num_atoms, p1, p2, distance):
    """Which constraints form rigid 3-site molecules, as OpenMM's CUDA platform decides [recalled]: atoms with
    exactly two constraints each, closing a triangle whose three atoms carry no other constraint; two of the
    three distances must be equal and the atom between them is the apex.  Returns (atoms (n,3), params (n,2)
    float32, number of constraints not covered)."""
    count = np.zeros(num_atoms, int)
    for a, b in zip(p1, p2):
        count[a] += 1
        count[b] += 1
    nbr = [dict() for _ in range(num_atoms)]
    for a, b, d in zip(p1, p2, distance):
        if count[a] == 2 and count[b] == 2:
            nbr[a][b] = d
            nbr[b][a] = d
    atoms, params = [], []
    for i in range(num_atoms):
        if len(nbr[i]) != 2:
            continue
        j, k = sorted(nbr[i])
        if len(nbr[j]) != 2 or len(nbr[k]) != 2 or k not in nbr[j]:
            continue
        if not (i < j and i < k):
            continue
        d12, d13, d23 = nbr[i][j], nbr[i][k], nbr[j][k]
        if d12 == d13:
            atoms.append((i, j, k)); params.append((d12, d23))
        elif d12 == d23:
            atoms.append((j, i, k)); params.append((d12, d13))
        elif d13 == d23:
            atoms.append((k, i, j)); params.append((d13, d12))
        else:
            raise ValueError("Two of the three distances constrained with SETTLE must be the same.")
    covered = 3 * len(atoms)
    return (np.asarray(atoms, np.int32).reshape(-1, 3), np.asarray(params, np.float32).reshape(-1, 2),
            len(p1) - covered)
