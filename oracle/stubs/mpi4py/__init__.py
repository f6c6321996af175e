"""Stand-in for mpi4py: the reference imports it but every call site passes use_mpi=False."""
from . import MPI
