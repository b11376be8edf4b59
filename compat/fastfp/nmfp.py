from fastfp_b200.nmfp import NMFP, CURN_container, GPEcorr_container, RN_container  # noqa: F401
