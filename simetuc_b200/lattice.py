"""Doped-lattice generation: the INPUT of the rate-equation path.

Host mirror of `simetuc/lattice.py:390-563` (`generate`): build the supercell, thin the
partially occupied sites, dope it with sensitizers (S, type 0) and activators (A, type 1),
and hand the ion table (positions, types, ground-state indices) to the ET build.

Differences from the reference, on purpose:
  * the reference draws from an unseeded global RNG (`lattice.py:93,144`), so its lattices are
    not reproducible; here every draw comes from one `numpy.random.Generator(PCG64(seed))`,
    in the reference's order (occupancy draws `:93`, then per kept site the pair
    "doped?" / "S or A?" `:150-162`);
  * the reference needs ASE for the supercell; the P-6 / P1-like cells used by the shipped
    configs have one position per listed site, so the supercell is written out directly in
    ASE's `repeat` order (`for m0: for m1: for m2: for site`);
  * the O(n^2) Python distance loop (`lattice.py:231-256`) is replaced by a chunked numpy
    minimum-image search here and by the CUDA pair build (`csrc/`, `stb_build_pairs`) on the
    device; the dense `dist_array` is only materialised on request.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .util import LatticeError


def cell_vectors(a, b, c, alpha, beta, gamma):
    """Rows = lattice vectors, ASE `cellpar_to_cell` convention: a1 along x, a2 in the xy plane
    (confirmed by the reference fixture positions)."""
    al, be, ga = np.deg2rad([alpha, beta, gamma])
    cos_a, cos_b, cos_g = (0.0 if abs(v - 90.0) < 1e-12 else np.cos(r)
                           for v, r in ((alpha, al), (beta, be), (gamma, ga)))
    sin_g = 1.0 if abs(gamma - 90.0) < 1e-12 else np.sin(ga)
    cy = (cos_a - cos_b * cos_g) / sin_g
    cz = np.sqrt(max(1.0 - cos_b ** 2 - cy ** 2, 0.0))
    return np.array([[a, 0.0, 0.0],
                     [b * cos_g, b * sin_g, 0.0],
                     [c * cos_b, c * cy, c * cz]], dtype=np.float64)


# Wyckoff positions of multiplicity 1 of P-6 (No. 174): listed sites on them are not multiplied by
# the symmetry operations, so the supercell can be written out without a space-group table.
_P6BAR_SINGLE = [(0.0, 0.0, 0.0), (0.0, 0.0, 0.5), (1 / 3, 2 / 3, 0.0), (1 / 3, 2 / 3, 0.5),
                 (2 / 3, 1 / 3, 0.0), (2 / 3, 1 / 3, 0.5)]


def _check_symmetry_supported(lat):
    """The reference expands the listed sites with `ase.spacegroup.crystal` (`lattice.py:80-107`) or
    reads a cif file; neither ASE nor a space-group table exists here.  Only lattices whose
    listed sites are the complete cell are accepted (P1, or P-6 with multiplicity-1 sites: every
    shipped `.cfg`); anything else raises instead of silently building the wrong crystal."""
    if lat.get('cif_file') or lat.get('ion_sites'):
        raise NotImplementedError('lattices defined by a cif file (`cif_file` / `ion_sites`) need the '
                                  'symmetry expansion of ASE, which this build does not have')
    sg = lat.get('spacegroup', 'P1')
    name = str(sg).replace(' ', '').upper()
    if name in ('P1', '1'):
        return
    if name in ('P-6', '174'):
        pos = lat['sites_pos']
        pos = pos if isinstance(pos[0], (list, tuple)) else [pos]
        for site in pos:
            frac = np.mod(np.asarray(site, dtype=np.float64), 1.0)
            if not any(np.allclose(frac, w, atol=1e-3) for w in _P6BAR_SINGLE):
                raise NotImplementedError('site %s of space group P-6 has multiplicity > 1: the '
                                          'symmetry expansion is not available' % (tuple(site),))
        return
    raise NotImplementedError('space group %r: only P1 and P-6 (multiplicity-1 sites) can be built '
                              'without a space-group table' % (sg,))


def check_lattice_settings(cte):
    """`lattice.py:40-78`: concentration / size / state-count validation."""
    lat = cte['lattice']
    S_conc, A_conc = float(lat['S_conc']), float(lat['A_conc'])
    if lat.get('radius') is None and int(lat['N_uc']) < 1:
        raise LatticeError('Wrong lattice settings.')
    if S_conc < 0 or A_conc < 0 or not 0 <= S_conc + A_conc <= 100.0:
        raise LatticeError('Wrong ion concentrations: %.2f%% Sensitizer, %.2f%% Activator. '
                           'Their sum must be between 0%% and 100%%.' % (S_conc, A_conc))
    _check_symmetry_supported(lat)
    pos, occ = lat['sites_pos'], lat['sites_occ']
    pos = pos if isinstance(pos[0], (list, tuple)) else [pos]
    occ = occ if isinstance(occ, (list, tuple)) else [occ]
    if len(pos) != len(occ):
        raise LatticeError('The number of sites must be the same in sites_pos and sites_occ.')
    nS, nA = cte['states']['sensitizer_states'], cte['states']['activator_states']
    if (S_conc != 0 and nS == 0) or (A_conc != 0 and nA == 0):
        raise LatticeError('The number of states of each ion cannot be zero.')
    return pos, occ


def mic_distances(pos_a, pos_b, cell, periodic=True, chunk=2048):
    """Minimum-image distances `[len(a), len(b)]` (float64).  Fractional wrap to the nearest
    image, then the 27 neighbouring images are searched (exact for the cells used here)."""
    pos_a = np.asarray(pos_a, dtype=np.float64)
    pos_b = np.asarray(pos_b, dtype=np.float64)
    out = np.empty((len(pos_a), len(pos_b)), dtype=np.float64)
    if not periodic:
        for s in range(0, len(pos_a), chunk):
            d = pos_b[None, :, :] - pos_a[s:s + chunk, None, :]
            out[s:s + chunk] = np.sqrt((d * d).sum(-1))
        return out
    inv = np.linalg.inv(cell)
    shifts = np.array([[i, j, k] for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                      dtype=np.float64) @ cell
    for s in range(0, len(pos_a), chunk):
        d = pos_b[None, :, :] - pos_a[s:s + chunk, None, :]
        frac = d @ inv
        d = (frac - np.rint(frac)) @ cell
        best = np.full(d.shape[:2], np.inf)
        for sh in shifts:
            e = d + sh
            np.minimum(best, (e * e).sum(-1), out=best)
        out[s:s + chunk] = np.sqrt(best)
    return out


@dataclass
class Lattice:
    """Ion table of one doped lattice realisation (what `lattice.generate` stores on disk,
    `lattice.py:527-548`, minus the O(n^2) lists which the device build regenerates)."""
    ion_type: np.ndarray                 # uint32 [n_ions], 0 = S, 1 = A (lattice order)
    positions: np.ndarray                # float64 [n_ions, 3], Angstrom
    cell: Optional[np.ndarray]           # float64 [3,3] supercell vectors (None = non-periodic)
    n_S_states: int
    n_A_states: int
    name: str = 'lattice'
    N_uc: int = 0
    S_conc: float = 0.0
    A_conc: float = 0.0
    radius: Optional[float] = None
    seed: Optional[int] = None
    dist_array: Optional[np.ndarray] = field(default=None, repr=False)

    @property
    def n_ions(self):
        return int(len(self.ion_type))

    @property
    def n_S(self):
        return int(np.count_nonzero(self.ion_type == 0))

    @property
    def n_A(self):
        return int(np.count_nonzero(self.ion_type != 0))

    @property
    def n_states(self):
        return self.n_S * self.n_S_states + self.n_A * self.n_A_states

    def ground_states(self):
        """`(index_S_i, index_A_j, initial_population)` - `lattice.py:259-298` semantics:
        int64 GS index per ion (-1 = other type), uint64 populations (1 at each GS)."""
        is_A = self.ion_type != 0
        width = np.where(is_A, self.n_A_states, self.n_S_states).astype(np.int64)
        start = np.cumsum(width) - width
        index_S_i = np.where(~is_A & (self.n_S_states != 0), start, -1).astype(np.int64)
        index_A_j = np.where(is_A & (self.n_A_states != 0), start, -1).astype(np.int64)
        pop = np.zeros(int(width.sum()), dtype=np.uint64)
        pop[start[width > 0]] = 1
        return index_S_i, index_A_j, pop

    def distances(self):
        """Dense MIC distance matrix (diagonal 0), cached."""
        if self.dist_array is None:
            d = mic_distances(self.positions, self.positions, self.cell,
                              periodic=self.cell is not None)
            d = np.triu(d, 1)                      # upper triangle mirrored: exactly symmetric
            self.dist_array = d + d.T
        return self.dist_array

    def lattice_info(self):
        return {'num_total': self.n_ions, 'num_activators': self.n_A,
                'num_sensitizers': self.n_S, 'energy_states': self.n_states,
                'sensitizer_states': self.n_S_states, 'activator_states': self.n_A_states}

    def save(self, path):
        np.savez_compressed(
            path, ion_type=self.ion_type, positions=self.positions,
            cell=np.zeros((0, 3)) if self.cell is None else self.cell,
            meta=np.array([self.n_S_states, self.n_A_states, self.N_uc,
                           -1 if self.seed is None else self.seed], dtype=np.int64),
            conc=np.array([self.S_conc, self.A_conc,
                           np.nan if self.radius is None else self.radius]),
            name=np.array(self.name),
            **({} if self.dist_array is None else {'dist_array': self.dist_array}))

    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as z:
            meta, conc = z['meta'], z['conc']
            return cls(ion_type=z['ion_type'], positions=z['positions'],
                       cell=None if z['cell'].shape[0] == 0 else z['cell'],
                       n_S_states=int(meta[0]), n_A_states=int(meta[1]), N_uc=int(meta[2]),
                       seed=None if meta[3] < 0 else int(meta[3]),
                       S_conc=float(conc[0]), A_conc=float(conc[1]),
                       radius=None if np.isnan(conc[2]) else float(conc[2]),
                       name=str(z['name']),
                       dist_array=z['dist_array'] if 'dist_array' in z.files else None)

    @classmethod
    def from_arrays(cls, ion_type, dist_array, n_S_states, n_A_states, positions=None, **kw):
        """Wrap reference-format arrays (e.g. a decoded fixture): distances are given."""
        ion_type = np.asarray(ion_type, dtype=np.uint32)
        if positions is None:
            positions = np.zeros((len(ion_type), 3))
        return cls(ion_type=ion_type, positions=np.asarray(positions, dtype=np.float64),
                   cell=None, n_S_states=int(n_S_states), n_A_states=int(n_A_states),
                   dist_array=np.asarray(dist_array, dtype=np.float64), **kw)


def generate(cte, seed=None, with_distances=False) -> Lattice:
    """Generate one doped lattice realisation from `cte.lattice` / `cte.states`
    (`lattice.py:390-563`).  `seed=None` draws a fresh seed from OS entropy (the reference's
    behaviour); pass an int for a reproducible lattice.  Raises `LatticeError` like the
    reference when no doped ion / no state results (`:464-467`)."""
    sites_pos, sites_occ = check_lattice_settings(cte)
    lat = cte['lattice']
    n_uc = int(lat['N_uc'])
    radius = lat.get('radius')
    S_conc, A_conc = float(lat['S_conc']) / 100.0, float(lat['A_conc']) / 100.0
    cellpar = lat.get('cell_par') or [lat[k] for k in ('a', 'b', 'c', 'alpha', 'beta', 'gamma')]
    unit = cell_vectors(*cellpar)
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0])
    rng = np.random.Generator(np.random.PCG64(seed))

    # supercell in ASE repeat order: m0 outermost, site innermost (lattice.py:87-89)
    n_sites = len(sites_pos)
    m = np.arange(n_uc)
    grid = np.stack(np.meshgrid(m, m, m, indexing='ij'), axis=-1).reshape(-1, 3)
    frac = (grid[:, None, :] + np.asarray(sites_pos, dtype=np.float64)[None, :, :]).reshape(-1, 3)
    pos = frac @ unit
    # thin partially occupied sites (lattice.py:92-94): keep if u <= occupancy
    occ = np.tile(np.asarray(sites_occ, dtype=np.float64), n_uc ** 3)
    pos = pos[rng.random(len(pos)) <= occ]
    cell = unit * n_uc
    if radius is not None:                       # nanoparticle (lattice.py:112-127)
        center = 0.5 * (pos.max(0) - pos.min(0)) + pos.min(0)
        pos = pos[np.linalg.norm(pos - center, axis=1) <= radius] - center
        cell = None
    # doping (lattice.py:150-162): two draws per site, in site order
    draws = rng.random((len(pos), 2))
    doped = draws[:, 0] < (S_conc + A_conc)
    is_S = draws[:, 1] * (S_conc + A_conc) < S_conc
    ion_type = np.where(is_S[doped], 0, 1).astype(np.uint32)
    pos = np.ascontiguousarray(pos[doped])
    nS_states, nA_states = cte['states']['sensitizer_states'], cte['states']['activator_states']
    out = Lattice(ion_type=ion_type, positions=pos, cell=cell, n_S_states=nS_states,
                  n_A_states=nA_states, name=lat.get('name', 'lattice'), N_uc=n_uc,
                  S_conc=float(lat['S_conc']), A_conc=float(lat['A_conc']), radius=radius,
                  seed=seed)
    if out.n_ions == 0 or out.n_states == 0:
        raise LatticeError('No doped ions generated: the lattice or the concentrations are '
                           'too small, or the number of energy states is zero!')
    if with_distances:
        out.distances()
    return out


def generate_arrays(cte, seed=None):
    """Dict view (`ion_type`, `dist_array`, `doped_lattice`, ...) of `generate` for tests."""
    lat = generate(cte, seed=seed, with_distances=True)
    iS, iA, pop = lat.ground_states()
    return {'ion_type': lat.ion_type, 'dist_array': lat.dist_array, 'doped_lattice': lat.positions,
            'indices_S_i': iS, 'indices_A_j': iA, 'initial_population': pop, 'lattice': lat}
