import json

import pytest


def test_ds_io_and_sweep(tmp_path):
    from deepspeed_b200.nvme import generate_aio_param, run_io_benchmark, run_sweep
    r = run_io_benchmark(str(tmp_path / "f.bin"), 1 << 20, read=False, loops=1)
    assert r["gb_per_s_max"] > 0
    r = run_io_benchmark(str(tmp_path / "f.bin"), 1 << 20, read=True, loops=1, validate=True)
    assert r["op"] == "read"
    res = run_sweep(str(tmp_path), "1M", {"block_size": ["64K", "256K"], "queue_depth": [4], "threads": [1, 2],
                                          "single_submit": [False], "overlap_events": [True]}, loops=1)
    best = generate_aio_param(res)
    assert set(best["aio"]) == {"block_size", "queue_depth", "intra_op_parallelism", "single_submit", "overlap_events"}


def test_launcher_resource_parsing(tmp_path):
    from deepspeed_b200.launcher import runner as R
    hf = tmp_path / "hostfile"
    hf.write_text("# comment\nworker-0 slots=4\nworker-1 slots=4\n")
    pool = R.fetch_hostfile(str(hf))
    assert list(pool.items()) == [("worker-0", 4), ("worker-1", 4)]
    act = R.parse_inclusion_exclusion(pool, "worker-0@worker-1:0,2", "")
    assert act == {"worker-0": [0, 1, 2, 3], "worker-1": [0, 2]}
    act = R.parse_inclusion_exclusion(pool, "", "worker-1:1,3")
    assert act["worker-1"] == [0, 2]
    act = R.parse_inclusion_exclusion(pool, "", "worker-0")
    assert list(act) == ["worker-1"]
    with pytest.raises(ValueError):
        R.parse_inclusion_exclusion(pool, "worker-9", "")
    with pytest.raises(ValueError):
        R._parse_hostfile(["worker-0 slots=4", "worker-0 slots=2"])
    enc = R.encode_world_info(act)
    assert R.decode_world_info(enc) == {"worker-1": [0, 1, 2, 3]}
    assert R.parse_num_nodes("2:4", True) == (2, 4)
    args = R.parse_args(["--num_gpus", "2", "--master_addr", "127.0.0.1", "train.py", "--foo", "1"])
    cmd = R.build_launch_cmd(args, enc)
    assert "deepspeed_b200.launcher.launch" in cmd and cmd[-3:] == ["train.py", "--foo", "1"]


def test_launch_env_and_multinode_cmds():
    from types import SimpleNamespace
    from deepspeed_b200.launcher import launch as L
    from deepspeed_b200.launcher.multinode_runner import RUNNERS
    a = L.parse_args(["--world_info", "x", "--node_rank", "1", "--master_addr", "10.0.0.1", "train.py", "--x"])
    launches = L.build_rank_env_and_cmds(a, {"h0": [0, 1], "h1": [0, 1, 2]})
    assert [e["RANK"] for e, _ in launches] == ["2", "3", "4"] and launches[0][0]["WORLD_SIZE"] == "5"
    assert launches[1][1][-2:] == ["--local_rank=1", "--x"]
    args = SimpleNamespace(user_script="t.py", user_args=["--a"], no_python=False, module=False, include="", exclude="",
                           launcher_args="", hostfile="/job/hostfile", master_addr="10.0.0.1", master_port=29500,
                           ssh_port=None, no_local_rank=False, save_pid=False, bind_cores_to_rank=False,
                           elastic_training=False, num_nodes=-1, num_gpus=-1)
    pool = {"h0": [0, 1], "h1": [0, 1]}
    for name, cls in RUNNERS.items():
        r = cls(args, "WORLD", pool)
        r.add_export("NCCL_DEBUG", "INFO")
        cmd = r.get_cmd({}, pool)
        assert "t.py" in cmd and any("NCCL_DEBUG" in str(c) for c in cmd), name


def test_env_report_runs(capsys):
    from deepspeed_b200.env_report import main
    main()
    out = capsys.readouterr().out
    assert "fused_adam" in out and "torch version" in out
