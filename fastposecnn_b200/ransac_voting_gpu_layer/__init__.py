"""Drop-in for the reference's lib/ransac_voting_gpu_layer package: ``ransac_voting`` mirrors the
pybind module (src/ransac_voting.cpp:102-107), ``ransac_voting_gpu`` the Python drivers."""
from . import ransac_voting  # noqa: F401
from . import ransac_voting_gpu  # noqa: F401
