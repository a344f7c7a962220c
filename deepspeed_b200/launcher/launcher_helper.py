"""MPI-launched entry: derive RANK/LOCAL_RANK/WORLD_SIZE from the MPI environment, then exec the user script
(reference ``launcher/launcher_helper.py``)."""
import argparse
import os
import subprocess
import sys


def env_mapping(env, rank_name_list=None, local_rank_name_list=None):
    def pick(names):
        vals = {env[n] for n in names if n in env}
        if len(vals) > 1:
            raise EnvironmentError(f"inconsistent rank variables {names}: {vals}")
        return vals.pop() if vals else None

    rank = pick(rank_name_list or ["PMIX_RANK", "PMI_RANK", "OMPI_COMM_WORLD_RANK", "MV2_COMM_WORLD_RANK", "SLURM_PROCID"])
    local = pick(local_rank_name_list or ["MPI_LOCALRANKID", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK",
                                          "SLURM_LOCALID"])
    if rank is None or local is None:
        raise EnvironmentError("no MPI/Slurm rank variables found in the environment")
    env["RANK"], env["LOCAL_RANK"] = rank, local
    return env


def parse_args(args=None):
    p = argparse.ArgumentParser(description="DeepSpeed launcher helper: maps the MPI/Slurm environment onto RANK/LOCAL_RANK "
                                "and runs the user script")
    p.add_argument("--launcher", default="mpich")
    p.add_argument("--module", action="store_true")
    p.add_argument("--no_python", action="store_true")
    p.add_argument("--no_local_rank", action="store_true")
    p.add_argument("user_script")
    p.add_argument("user_args", nargs=argparse.REMAINDER)
    return p.parse_args(args)


def main(argv=None):
    a = parse_args(argv)
    env = env_mapping(os.environ.copy())
    cmd = ([] if a.no_python else [sys.executable, "-u"] + (["-m"] if a.module else [])) + [a.user_script]
    if not a.no_local_rank:
        cmd.append(f"--local_rank={env['LOCAL_RANK']}")
    sys.exit(subprocess.call(cmd + a.user_args, env=env))


if __name__ == "__main__":
    main()
