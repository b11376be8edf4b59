"""Physical constants used on the Fp hot path.

Mirrors the values of the reference's ``fastfp/constants.py:7-9`` (which takes them from
``scipy.constants``): a Julian year in seconds, a day in seconds and ``fyr = 1/yr``.
They are restated as literals so the package does not need scipy at import time.
"""

yr = 31557600.0  # scipy.constants.Julian_year  (365.25 d)
day = 86400.0  # scipy.constants.day
fyr = 1.0 / yr
