"""``MpiBackend`` name for the compressed all-reduce (reference ``runtime/comm/mpi.py`` uses mpi4py + cupy).  There is
no MPI dependency here: the same algorithm runs over the torch.distributed group, which an ``mpirun`` launch initialises
through ``comm.init_distributed(auto_mpi_discovery=True)``."""
from .compressed import CompressedBackend


class MpiBackend(CompressedBackend):

    def __init__(self, cuda_aware=True, mpu=None, group=None):
        super().__init__(mpu=mpu, group=group)
        self.cuda_aware = cuda_aware
