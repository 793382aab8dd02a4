"""Host-side pieces of bench.py that run without a GPU: the ncu-summary parser behind `roofline.traffic`, the per-workload
config dictionary the reference arm must reproduce verbatim, and the evaluation byte count."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def test_ncu_traffic_parses_committed_summaries():
    big = bench.ncu_traffic(bench.NCU_TRAFFIC["train_b65536"])
    small = bench.ncu_traffic(bench.NCU_TRAFFIC["train_b64"])
    # one batch-65536 step moves 5-10 GB through HBM, one batch-64 step ~0.1-0.2 GB (the Adam streams)
    assert big is not None and 5e9 < big < 1e10
    assert small is not None and 1e8 < small < 3e8
    assert bench.ncu_traffic("no_such_file.csv") is None


def test_ncu_traffic_skips_comment_lines(tmp_path, monkeypatch):
    d = tmp_path / "profiles"
    d.mkdir()
    (d / "x.csv").write_text("dram__bytes_read.sum,Gbyte,1.5\n# a comment without commas\n# one, with, commas\n"
                             "dram__bytes_write.sum,Mbyte,500\ngpu__time_duration.sum,usecond,10\n")
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert abs(bench.ncu_traffic("x.csv") - 2.0e9) < 1.0


def test_workload_config_is_the_same_for_both_arms():
    # the driver compares the `config` of the two arms: nothing arm-specific may live inside it
    a = bench.workload_config("train_b65536", bench.LARGE_BATCH, 1, True)
    b = bench.workload_config("train_b65536", bench.LARGE_BATCH, 1, True)
    assert a == b and "reference_arm" not in a and "impl" not in a
    assert a["batch_per_gpu"] == 65536 and a["global_batch"] == 65536
    assert bench.workload_config("train_b65536", bench.LARGE_BATCH, 8, True)["global_batch"] == 8 * 65536
    json.dumps(a)


def test_constants_follow_the_survey():
    assert bench.EVAL_BYTES_PER_POSE == 2 * 16 * 3 * 4
    assert bench.batch_of("train_b64") == 64 and bench.batch_of("train_b65536") == 65536
