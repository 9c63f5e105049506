"""Domain types for the rate-equation path: ions, transitions, excitations, ET processes.

Host-side mirror of the reference's `simetuc/util.py:82-292` (same class names, constructor
arguments and attribute meaning, including the reference's spelling
`EneryTransferProcess`) so that a `cte` settings dictionary written for the reference can be
handed to this package unchanged.  Only what the hot path reads is implemented.
"""
import contextlib
import logging
from collections import namedtuple
from enum import Enum
from typing import Optional, Sequence

Conc = namedtuple('Concentration', ['S_conc', 'A_conc'])


class LatticeError(Exception):
    """Wrong lattice / state-count settings (reference: `lattice.py:34`)."""


class IonType(Enum):
    sensitizer = 0
    activator = 1
    S = 0
    A = 1


def ion_code(ion) -> int:
    """0 for a sensitizer, 1 for an activator; accepts this package's or the reference's enum."""
    return int(getattr(ion, 'value', ion))


class Transition:
    """`ion(state_i -> state_f)` (reference: `util.py:90-127`)."""

    def __init__(self, ion, state_i: int, state_f: int,
                 label_ion: str = '', label_i: str = '', label_f: str = '') -> None:
        self.ion = ion
        self.state_i = state_i
        self.state_f = state_f
        self.label_ion, self.label_i, self.label_f = label_ion, label_i, label_f

    def _key(self):
        return (ion_code(self.ion), self.state_i, self.state_f)

    def __eq__(self, other):
        if not isinstance(other, Transition):
            return NotImplemented
        return self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        ion = self.label_ion or 'SA'[ion_code(self.ion)]
        return '%s(%s: %s->%s)' % (type(self).__name__, ion,
                                   self.label_i or self.state_i, self.label_f or self.state_f)


class DecayTransition(Transition):
    """A decay rate (s^-1) of `state_i`, or a branching ratio `state_i -> state_f`
    (reference: `util.py:130-200`).  `.value` is what the optimiser varies."""

    def __init__(self, ion, state_i: int, state_f: int,
                 decay_rate: Optional[float] = None, branching_ratio: Optional[float] = None,
                 label_ion: str = '', label_i: str = '', label_f: str = '') -> None:
        super().__init__(ion, state_i, state_f, label_ion, label_i, label_f)
        if decay_rate is None and branching_ratio is None:
            raise ValueError('Either the decay rate or the branching ratio must be given.')
        self.decay_rate = decay_rate
        self.branching_ratio = branching_ratio
        self.name = ('k_%d%d' if decay_rate is not None else 'B_%d%d') % (state_i, state_f)

    def __eq__(self, other):
        if isinstance(other, DecayTransition):
            return (self._key() == other._key() and self.decay_rate == other.decay_rate
                    and self.branching_ratio == other.branching_ratio)
        if isinstance(other, Transition):
            return self._key() == other._key()
        return NotImplemented

    __hash__ = Transition.__hash__

    @property
    def value(self):
        return self.decay_rate if self.decay_rate is not None else self.branching_ratio

    @value.setter
    def value(self, val):
        if self.decay_rate is not None:
            self.decay_rate = val
        else:
            self.branching_ratio = val


class Excitation:
    """One absorption step of an excitation source (reference: `util.py:203-232`)."""

    def __init__(self, ion, state_i: int, state_f: int, active: bool, degeneracy: float,
                 pump_rate: float, power_dens: float, t_pulse: Optional[float] = None,
                 label_ion: str = '', label_i: str = '', label_f: str = '') -> None:
        self.transition = Transition(ion, state_i, state_f, label_ion, label_i, label_f)
        self.active = active
        self.degeneracy = degeneracy
        self.pump_rate = pump_rate
        self.power_dens = power_dens
        self.t_pulse = t_pulse

    def __eq__(self, other):
        if not isinstance(other, Excitation):
            return NotImplemented
        return all(getattr(self, a) == getattr(other, a) for a in
                   ('transition', 'active', 'degeneracy', 'pump_rate', 'power_dens', 't_pulse'))

    def __repr__(self):
        return repr(self.transition).replace('Transition', 'Excitation')[:-1] + \
            (', active)' if self.active else ', inactive)')


class EneryTransferProcess:
    """Pair (2 transitions) or cooperative (3 transitions) energy-transfer process.

    `indices` lists the initial states of every transition, then the final ones
    (`(ii, fi, if, ff)` for a pair process); `type` is the tuple of ion types
    `(donor, partner)` (reference: `util.py:235-292`).
    """

    def __init__(self, transitions: Sequence[Transition], mult: int, strength: float,
                 strength_avg: Optional[float] = None, name: str = '') -> None:
        self.transitions = tuple(transitions)
        self.mult = int(mult)
        self.strength = strength
        self._strength_avg = strength_avg
        self.name = name
        self.indices = tuple([t.state_i for t in self.transitions] +
                             [t.state_f for t in self.transitions])
        self.type = tuple(t.ion for t in self.transitions)
        self.cooperative = len(self.transitions) == 3

    @property
    def strength_avg(self):
        return self.strength if self._strength_avg is None else self._strength_avg

    @property
    def value(self):
        return self.strength

    @value.setter
    def value(self, val):
        self.strength = val

    def __eq__(self, other):
        if not isinstance(other, EneryTransferProcess):
            return NotImplemented
        return self.transitions == other.transitions and self.mult == other.mult

    def __repr__(self):
        ini = '+'.join('%s(%s)' % ('SA'[ion_code(t.ion)], t.state_i) for t in self.transitions)
        fin = '+'.join('%s(%s)' % ('SA'[ion_code(t.ion)], t.state_f) for t in self.transitions)
        return 'ETProcess(%s->%s, n=%d, C=%.2g)' % (ini, fin, self.mult, self.strength)


EnergyTransferProcess = EneryTransferProcess


@contextlib.contextmanager
def disable_loggers(names):
    """Silence the named loggers inside the block (reference: `util.py:354-368`)."""
    saved = []
    for name in names:
        logger = logging.getLogger(name)
        saved.append((logger, logger.disabled))
        logger.disabled = True
    try:
        yield
    finally:
        for logger, state in saved:
            logger.disabled = state
