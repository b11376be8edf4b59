from fastfp_b200.constants import day, fyr, yr  # noqa: F401
