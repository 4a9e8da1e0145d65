"""Numerical contract of the hot path: the reference's constants
(spectral_simulator/constants.py:2-7), written as the same *expressions* so the
IEEE-754 values are identical (2.998 * 10**10 == 29980000000.000004).  They are
not CODATA values and must not be corrected.  The CUDA side carries the same
doubles as hex literals (csrc/lte_common.cuh); tests/test_host.py::test_constants_match_cuda_header_bit_for_bit checks
the two agree bit for bit."""
kcm = 0.69503476          # cm^-1 / K   (functions.py:323)
ckm = 2.998 * 10**5       # km/s
ccm = 2.998 * 10**10      # cm/s
cm = 2.998 * 10**8        # m/s
h = 6.626 * 10**(-34)     # J s
k = 1.381 * 10**(-23)     # J/K

TBG = 2.7                 # inference.py:57
CT = 300                  # classes.py:19
FWHM_TO_SIGMA = 2.355     # inference.py:53

# ANSI colours used by the reference's console messages (constants.py:10-14)
CYAN = "\033[36m"
GRAY = "\033[90m"
RED = "\033[31m"
GREEN = "\033[92m"
RESET = "\033[0m"
