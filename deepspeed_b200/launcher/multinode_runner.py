"""Multi-node fan-out backends.  Reference: ``launcher/multinode_runner.py`` (PDSH :56, OpenMPI :125, MPICH :191,
IMPI :266, Slurm :355, MVAPICH :407).  Each runner only *builds a command line*; the runner module executes it."""
import os
import shutil
import sys
from abc import ABC, abstractmethod
from shlex import quote

from . import constants as K


class MultiNodeRunner(ABC):

    def __init__(self, args, world_info_base64, resource_pool=None):
        self.args = args
        self.validate_args()  # (reference: the constructor validates, multinode_runner.py:21)
        self.world_info_base64 = world_info_base64
        self.resource_pool = resource_pool or {}
        self.exports = {}
        self.user_arguments = self.parse_user_args()
        self.user_script = args.user_script

    @staticmethod
    def _slots(v):
        """Slot count of a resource-pool entry: the pool maps host -> slot count (hostfile form) or host -> slot list."""
        return len(v) if hasattr(v, "__len__") else int(v)

    def parse_user_args(self):
        """User-script arguments as they must appear on the launcher's command line (runners that go through a remote shell
        override this to quote arguments containing spaces)."""
        return list(self.args.user_args)

    def validate_args(self):
        """Reject argument combinations this runner cannot honour (overridden per backend)."""
        if getattr(self.args, "include", "") and getattr(self.args, "exclude", ""):
            raise ValueError("--include and --exclude are mutually exclusive")

    @abstractmethod
    def backend_exists(self):
        ...

    @abstractmethod
    def get_cmd(self, environment, active_resources):
        ...

    def add_export(self, key, var):
        self.exports[key.strip()] = f'"{var.strip()}"' if " " in var.strip() else var.strip()

    @property
    def name(self):
        return self.__class__.__name__.replace("Runner", "").lower()  # "pdsh", "openmpi", "mpich", "impi", "slurm", "mvapich"

    def _launch_tail(self):
        tail = []
        if not self.args.no_python:
            tail += [sys.executable, "-u"]
            if self.args.module:
                tail.append("-m")
        return tail + [self.user_script] + self.user_arguments

    def _total_procs(self):
        return sum(self._slots(v) for v in self.resource_pool.values())

    def _per_node(self):
        return self._slots(next(iter(self.resource_pool.values()))) if self.resource_pool else 1


class PDSHRunner(MultiNodeRunner):

    def backend_exists(self):
        return shutil.which("pdsh") is not None

    def get_cmd(self, environment, active_resources):
        environment["PDSH_RCMD_TYPE"] = "ssh"
        if self.args.ssh_port is not None:
            environment["PDSH_SSH_ARGS_APPEND"] = f"{environment.get('PDSH_SSH_ARGS_APPEND', '')} -p {self.args.ssh_port}"
        workers = ",".join(active_resources.keys())
        exports = "".join(f"export {k}={quote(v)}; " for k, v in self.exports.items())
        launch = [exports + f"cd {os.path.abspath('.')};", sys.executable, "-u", "-m", "deepspeed_b200.launcher.launch",
                  f"--world_info={self.world_info_base64}", "--node_rank=%n", f"--master_addr={self.args.master_addr}",
                  f"--master_port={self.args.master_port}"]
        for flag in ("no_python", "module", "no_local_rank", "save_pid", "bind_cores_to_rank"):
            if getattr(self.args, flag):
                launch.append(f"--{flag}")
        if self.args.elastic_training:
            launch += ["--enable_elastic_training", f"--max_elastic_nodes={self.args.max_elastic_nodes}",
                       f"--min_elastic_nodes={self.args.min_elastic_nodes}"]
        extra = self.args.launcher_args.split() if self.args.launcher_args else []
        pdsh = ["pdsh", "-S", "-f", str(K.PDSH_MAX_FAN_OUT), "-w", workers] + extra
        # what the runner executes on the workers when the job is interrupted: stop every per-node launcher of this job
        kill_cmd = pdsh + ["pkill", "-f", "deepspeed_b200.launcher.launch"]
        return pdsh + launch + [self.user_script] + self.user_arguments, kill_cmd, environment


class OpenMPIRunner(MultiNodeRunner):

    def __init__(self, args, world_info_base64, resource_pool=None):
        super().__init__(args, world_info_base64, resource_pool)
        self.add_export("UCX_TLS", "tcp")

    def backend_exists(self):
        return shutil.which("ompi_info") is not None

    def validate_args(self):
        super().validate_args()
        self._setup_mpi_environment()
        if getattr(self.args, "include", "") or getattr(self.args, "exclude", ""):
            raise ValueError(f"{self.name} backend does not support worker include/exclusion")
        if getattr(self.args, "num_nodes", -1) != -1 or getattr(self.args, "num_gpus", -1) != -1:
            raise ValueError(f"{self.name} backend does not support limiting num nodes/gpus")

    def _setup_mpi_environment(self):
        """The OpenMPI path is entered from inside an MPI job: mirror its rank variables into the names the engine reads
        (reference ``multinode_runner.py:147``)."""
        required = ["OMPI_COMM_WORLD_LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE"]
        if not all(v in os.environ for v in required):
            raise EnvironmentError("MPI environment variables are not set. Ensure you are running the script with an "
                                   "MPI-compatible launcher.")
        os.environ["LOCAL_RANK"] = os.environ["OMPI_COMM_WORLD_LOCAL_RANK"]
        os.environ["RANK"] = os.environ["OMPI_COMM_WORLD_RANK"]
        os.environ["WORLD_SIZE"] = os.environ["OMPI_COMM_WORLD_SIZE"]

    def get_cmd(self, environment, active_resources):
        from shlex import split
        launcher_args = split(self.args.launcher_args) if self.args.launcher_args else []
        # `--mca btl_tcp_if_include eth0` unless the user chose an interface through --launcher_args
        btl = ["--mca", "btl_tcp_if_include", "eth0"]
        for i in range(len(launcher_args) - 1):
            if launcher_args[i] in ("-mca", "--mca") and launcher_args[i + 1] == "btl_tcp_if_include":
                btl = []
                break
        cmd = ["mpirun", "-n", str(self._total_procs()), "-hostfile", str(self.args.hostfile), "--mca", "btl", "^openib"] + \
            btl + launcher_args
        for k, v in self.exports.items():
            cmd += ["-x", f"{k}={v}"]
        return cmd + self._launch_tail()


class MPICHRunner(MultiNodeRunner):

    def backend_exists(self):
        return shutil.which("mpirun") is not None

    def get_cmd(self, environment, active_resources):
        if self.args.include or self.args.exclude:
            raise ValueError(f"{self.name} backend does not support worker include/exclusion")
        hosts = list(active_resources.keys())
        cmd = ["mpirun"] + (self.args.launcher_args.split() if self.args.launcher_args else [])
        first = True
        rank = 0
        total = self._total_procs()
        for h in hosts:
            for lr in range(self._slots(active_resources[h])):
                if not first:
                    cmd.append(":")
                first = False
                cmd += ["-n", "1", "-host", h, "-env", "RANK", str(rank), "-env", "LOCAL_RANK", str(lr), "-env",
                        "WORLD_SIZE", str(total), "-env", "LOCAL_SIZE", str(self._slots(active_resources[h])), "-env",
                        "MASTER_ADDR", str(self.args.master_addr), "-env", "MASTER_PORT", str(self.args.master_port)]
                for k, v in self.exports.items():
                    cmd += ["-env", k, v]
                cmd += self._launch_tail()
                rank += 1
        return cmd


class IMPIRunner(MPICHRunner):

    def backend_exists(self):
        return shutil.which("mpiexec.hydra") is not None or shutil.which("mpirun") is not None

    def get_cmd(self, environment, active_resources):
        cmd = super().get_cmd(environment, active_resources)
        return ["mpirun", "-ppn", str(self._per_node())] + cmd[1:]


class SlurmRunner(MultiNodeRunner):

    def backend_exists(self):
        return shutil.which("sinfo") is not None

    def get_cmd(self, environment, active_resources):
        assert not getattr(self.args, "detect_nvlink_pairs", False), "slurm backend does not support remapping visible devices"
        cmd = ["srun", "-n", str(self._total_procs())] + (self.args.launcher_args.split() if self.args.launcher_args else [])
        if getattr(self.args, "comment", ""):
            cmd += ["--comment", self.args.comment]
        if self.args.include:
            cmd += ["--include", self.args.include]
        if self.args.exclude:
            cmd += ["--exclude", self.args.exclude]
        if self.args.num_nodes > 0:
            cmd += ["--nodes", str(self.args.num_nodes)]
        if self.args.num_gpus > 0:
            cmd += ["--gpus", str(self.args.num_gpus)]
        exports = "--export=ALL" + "".join(f",{k}={v}" for k, v in self.exports.items())
        return cmd + [exports] + self._launch_tail()


class MVAPICHRunner(MultiNodeRunner):

    def backend_exists(self):
        return shutil.which("mpiname") is not None

    def get_cmd(self, environment, active_resources):
        if self.args.include or self.args.exclude:
            raise ValueError(f"{self.name} backend does not support worker include/exclusion")
        with open(K.MVAPICH_TMP_HOSTFILE, "w") as f:
            for h in active_resources:
                f.write(f"{h}\n")
        self.add_export("MV2_SMP_USE_CMA", "0")
        self.add_export("MV2_DEBUG_SHOW_BACKTRACE", "1")
        self.add_export("MV2_USE_CUDA", "1")
        self.add_export("MV2_SUPPORT_DL", "1")
        self.add_export("MV2_ENABLE_AFFINITY", "0")
        cmd = ["mpirun", "-np", str(self._total_procs()), "-ppn", str(self._per_node()), "--hostfile", K.MVAPICH_TMP_HOSTFILE]
        cmd += self.args.launcher_args.split() if self.args.launcher_args else []
        for k, v in self.exports.items():
            cmd += ["-env", f"{k}={v}"]
        return cmd + self._launch_tail()


RUNNERS = {K.PDSH_LAUNCHER: PDSHRunner, K.OPENMPI_LAUNCHER: OpenMPIRunner, K.MPICH_LAUNCHER: MPICHRunner,
           K.IMPI_LAUNCHER: IMPIRunner, K.SLURM_LAUNCHER: SlurmRunner, K.MVAPICH_LAUNCHER: MVAPICHRunner}
