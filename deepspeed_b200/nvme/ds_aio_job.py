"""Shell job wrapper used by the sweep (reference ``nvme/ds_aio_job.py``)."""
import contextlib
import subprocess
from dataclasses import dataclass
from typing import List, Optional


@dataclass
class Job:
    cmd_line: List[str]
    output_file: Optional[str] = None
    work_dir: Optional[str] = None
    output_fd = None

    def cmd(self):
        return self.cmd_line

    def get_stdout(self):
        return self.output_fd

    get_stderr = get_stdout

    def get_cwd(self):
        return self.work_dir

    @contextlib.contextmanager
    def redirected(self):
        """Open the log file for the duration of the run (stdout and stderr share it)."""
        self.output_fd = open(self.output_file, "w") if self.output_file else None
        try:
            yield self.output_fd
        finally:
            if self.output_fd is not None:
                self.output_fd.close()
                self.output_fd = None

    # reference-style explicit open / close
    def open_output_file(self):
        if self.output_file:
            self.output_fd = open(self.output_file, "w")

    def close_output_file(self):
        if self.output_fd is not None:
            self.output_fd.close()
            self.output_fd = None


def run_job(job, verbose=False):
    line = " ".join(job.cmd())
    if verbose:
        print(f"args = {line}")
    with job.redirected() as fd:
        rc = subprocess.run(line, shell=True, stdout=fd, stderr=fd, cwd=job.get_cwd()).returncode
    assert rc == 0, f"'{line}' failed with exit code {rc}"
    return rc
