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
    import os
    for k, v in (("OMPI_COMM_WORLD_LOCAL_RANK", "0"), ("OMPI_COMM_WORLD_RANK", "0"), ("OMPI_COMM_WORLD_SIZE", "1")):
        os.environ[k] = v  # the OpenMPI runner validates that it runs inside an MPI job (reference behaviour)
    try:
        for pool in ({"h0": [0, 1], "h1": [0, 1]}, {"h0": 2, "h1": 2}):  # slot lists (runner) and slot counts (hostfile form)
            for name, cls in RUNNERS.items():
                r = cls(args, "WORLD", pool)
                assert r.name == name or name.startswith(r.name), (r.name, name)
                r.add_export("NCCL_DEBUG", "INFO")
                got = r.get_cmd({}, pool)
                cmd = got[0] if isinstance(got, tuple) else got  # pdsh: (cmd, kill_cmd, env)
                if isinstance(got, tuple):
                    assert got[1][0] == "pdsh" and got[2]["PDSH_RCMD_TYPE"] == "ssh"
                assert "t.py" in cmd and any("NCCL_DEBUG" in str(c) for c in cmd), name
    finally:
        for k in ("OMPI_COMM_WORLD_LOCAL_RANK", "OMPI_COMM_WORLD_RANK", "OMPI_COMM_WORLD_SIZE", "LOCAL_RANK", "RANK",
                  "WORLD_SIZE"):
            os.environ.pop(k, None)


def test_env_report_runs(capsys):
    from deepspeed_b200.env_report import main
    main()
    out = capsys.readouterr().out
    assert "fused_adam" in out and "torch version" in out


def test_nvme_tooling_sweep_logs_and_param_generation(tmp_path):
    """reference-dialect ds_io args, a 2-point sweep with config-encoded log names, log parsing, best-config selection."""
    import pytest
    from deepspeed_b200.nvme import ds_aio_args, parse_nvme_stats, perf_generate_param, perf_run_sweep
    from deepspeed_b200.nvme.ds_aio_job import Job, run_job
    from deepspeed_b200.nvme.test_ds_aio import ds_io_main
    from deepspeed_b200.nvme.test_ds_aio_utils import get_block_size_and_count, refine_integer_value
    from deepspeed_b200.nvme.validate_async_io import main as validate
    assert refine_integer_value("4K") == 4096 and refine_integer_value("2m") == 2 << 20 and refine_integer_value("17") == 17
    assert get_block_size_and_count(3 << 20) == (1 << 20, 3)
    io = tmp_path / "io"
    io.mkdir()
    args = ds_aio_args.get_validated_args(["--folder", str(io), "--io_size", "1M", "--block_size", "128K", "--loops", "1", "--read"])
    assert args.io_size == 1 << 20 and args.mapping_list == [(0, str(io))]
    with pytest.raises(SystemExit):
        ds_aio_args.get_validated_args(["--io_size", "1M"])
    assert ds_io_main(["--folder", str(io), "--io_size", "1M", "--block_size", "128K", "--loops", "1"]) > 0  # write
    assert ds_io_main(["--folder", str(io), "--io_size", "1M", "--block_size", "128K", "--loops", "1", "--read"]) > 0
    cfg = tmp_path / "sweep.json"
    cfg.write_text('{"block_size": ["128K", "256K"], "queue_depth": [4], "io_parallel": [1], "single_submit": [false]}')
    sargs = perf_run_sweep.parse_sweep_arguments(["--nvme_dir", str(io), "--sweep_config", str(cfg), "--io_size", "1M", "--log_dir",
                                                  str(tmp_path / "logs")])
    assert perf_run_sweep.validate_arguments(sargs)
    log_dir = perf_run_sweep.sweep_main(sargs)
    rlogs = sorted(p.name for p in (tmp_path / "logs" / "_aio_bench_read_logs").iterdir())
    assert rlogs == ["read_block_overlap_t1_p1_d4_bs128K.txt", "read_block_overlap_t1_p1_d4_bs256K.txt"]
    res, keys = parse_nvme_stats.get_sorted_results(str(tmp_path / "logs" / "_aio_bench_read_logs"), "read_speed")
    assert len(res) == 2 and all(v > 0 for v in res.values())
    assert parse_nvme_stats.get_thread_count("x/read_block_overlap_t8_p2_d4_bs1M.txt") == 16
    param = perf_generate_param.generate_main(log_dir)
    assert param["queue_depth"] == 4 and param["block_size"] in (128 << 10, 256 << 10) and param["single_submit"] == "false"
    out = tmp_path / "job.txt"
    run_job(Job(["echo", "hello"], str(out)))
    assert out.read_text().strip() == "hello"
    assert validate() is True


def test_autotuning_scheduler_reservations_and_concurrency(tmp_path):
    import json
    import os
    import threading
    import time
    import types
    from deepspeed_b200.autotuning import scheduler as S
    n = S.Node("h", 4)
    a = n.reserve_slots(3)
    assert a == [0, 1, 2] and n.reserve_slots(2) is None
    r = S.Reservation(n, a)
    assert S.include_string([r]) == "h:0,1,2"
    r.restore_slots()
    assert n.idle_slots == [0, 1, 2, 3]
    assert S.get_user() and S.get_job_id()
    live, peak, lock = [0], [0], threading.Lock()

    def runner(exp, rd):
        with lock:
            live[0] += 1
            peak[0] = max(peak[0], live[0])
        time.sleep(0.15)
        with open(os.path.join(rd, "metrics.json"), "w") as f:
            json.dump({"throughput": exp["num_gpus"] * 10.0 + exp["exp_id"]}, f)
        with lock:
            live[0] -= 1

    rm = S.ResourceManager(types.SimpleNamespace(user_script="x.py", user_args=[]), ["localhost"], 4, str(tmp_path), None,
                           runner=runner)
    exps = [{"name": f"e{i}", "ds_config": {}, "num_gpus": g, "num_nodes": 1} for i, g in enumerate([2, 2, 4, 1, 8])]
    rm.schedule_experiments_dicts(exps)
    rm.run()
    assert peak[0] == 2  # the two 2-GPU experiments ran side by side; the 4-GPU one had to wait for both
    assert len(rm.finished) == 5 and "needs 1x8" in rm.finished[4][1]
    best, val = rm.parse_results("throughput")
    assert best["name"] == "e2" and val == 42.0
    assert rm.status() == "localhost (4 idle gpus)"


def test_autotuner_micro_batch_search_helpers(tmp_path):
    import json
    import os
    import types
    from deepspeed_b200.autotuning.autotuner import Autotuner
    cfg_path = tmp_path / "ds.json"
    cfg_path.write_text(json.dumps({"train_micro_batch_size_per_gpu": "auto", "gradient_accumulation_steps": "auto",
                                    "fp16": {"enabled": True},
                                    "autotuning": {"enabled": True, "results_dir": str(tmp_path / "res"),
                                                   "exps_dir": str(tmp_path / "exps"), "max_train_batch_size": 64,
                                                   "arg_mappings": {"gradient_accumulation_steps": "--gas"}}}))
    limit = 11  # the fake runner "runs out of memory" above this micro batch

    def runner(exp, rd):
        mbs = exp["ds_config"]["train_micro_batch_size_per_gpu"]
        if mbs <= limit:
            with open(os.path.join(rd, "metrics.json"), "w") as f:
                json.dump({"throughput": 100.0 * mbs / (mbs + 4)}, f)

    args = types.SimpleNamespace(user_script="train.py", user_args=["--deepspeed_config", str(cfg_path), "--gas", "2"],
                                 hostfile=None)
    at = Autotuner(args, {"localhost": [0, 1]}, runner=runner)
    assert at.fp16_enabled() and at.get_gas_from_user_config() == 2 and at.get_val_from_user_args("gradient_accumulation_steps") == "2"
    assert at.max_train_micro_batch_size_per_gpu() <= 32 and at.get_activation_memory_per_gpu() is None
    assert at.get_min_max_micro_batch_size(1, 1, 16) == (1, 11)  # binary search finds the memory limit
    vals, max_tbs = at.get_tuning_micro_batch_size_list(1, 11, 3)
    assert vals == [1, 6, 11] and max_tbs == 11 * 2 * 2
    assert at.get_tuning_micro_batch_size_list(0, 4, 3) == ([], 0)
    best = at.run_tuning_micro_batch_sizes(vals, None, 1, 4, stage=1)
    assert best == 11 and at.get_plateau_mbs("z1") >= 1
    assert at.run_ds_config({"train_micro_batch_size_per_gpu": 50}, "too_big") is None
