from fastfp_b200.utils import get_mats_fp, get_mats_nmfp, get_xCy, initialize_pta  # noqa: F401
