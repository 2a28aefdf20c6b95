"""Combination rules and the ideal titration curve (reference: src/modules/utils.py:2-16).

Same call signatures as the reference.  Two of its defects are not reproduced: an unknown rule
name raises ValueError here (the reference *returns* the exception object), and the rule names
are accepted either way round (the reference labels the arithmetic mean of sigma "Berthelot" and
the geometric mean of epsilon "Lorentz", the opposite of the textbook naming).
"""
_RULE_NAMES = ("Lorentz", "Berthelot", "Lorentz-Berthelot")


def _check_rule(rule):
    if rule not in _RULE_NAMES:
        raise ValueError(f"No combination rule defined for {rule!r}")


def combination_rule_epsilon(rule, eps1, eps2):
    """Geometric mean of the two well depths."""
    _check_rule(rule)
    return (eps1 * eps2) ** 0.5


def combination_rule_sigma(rule, sig1, sig2):
    """Arithmetic mean of the two diameters."""
    _check_rule(rule)
    return 0.5 * (sig1 + sig2)


def ideal_alpha(pH, pK):
    """Degree of deprotonation of an ideal (non-interacting) weak site: 1/(1 + 10^(pK - pH))."""
    return 1.0 / (1.0 + 10.0 ** (pK - pH))
