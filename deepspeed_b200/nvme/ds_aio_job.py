"""Shell job wrapper used by the sweep (reference ``nvme/ds_aio_job.py``)."""
import subprocess


class Job:

    def __init__(self, cmd_line, output_file=None, work_dir=None):
        self.cmd_line, self.output_file, self.work_dir = cmd_line, output_file, work_dir
        self.output_fd = None

    def cmd(self):
        return self.cmd_line

    def get_stdout(self):
        return self.output_fd

    get_stderr = get_stdout

    def get_cwd(self):
        return self.work_dir

    def open_output_file(self):
        if self.output_file is not None:
            self.output_fd = open(self.output_file, "w")

    def close_output_file(self):
        if self.output_fd is not None:
            self.output_fd.close()
            self.output_fd = None


def run_job(job, verbose=False):
    args = " ".join(job.cmd())
    if verbose:
        print(f"args = {args}")
    job.open_output_file()
    try:
        proc = subprocess.run(args=args, shell=True, stdout=job.get_stdout(), stderr=job.get_stderr(), cwd=job.get_cwd())
    finally:
        job.close_output_file()
    assert proc.returncode == 0, f"'{args}' failed with exit code {proc.returncode}"
    return proc.returncode
