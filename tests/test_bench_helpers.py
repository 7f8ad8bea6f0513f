"""CPU: the arithmetic bench.py reports with -- algorithmic bytes and the parsing of
the nvidia-smi clock samples."""
import time

import bench as B


def test_algorithmic_bytes_matches_survey_formula():
    # identity features, d = 64: R = 260; forward = 20 + 2*276 + S*276 + update-test 40; update = 3*780
    c = {"positives": 1, "negatives_drawn": 1, "updates": 1}
    assert B.algorithmic_bytes(c, 64) == 20 + 2 * 276 + 276 + 40 + 2340
    c = {"positives": 10, "negatives_drawn": 100, "updates": 0}   # S = 10, no update
    assert B.algorithmic_bytes(c, 64) == 10 * (20 + 552) + 100 * 276
    # survey's C2 steady state (S = 2.7, U = 0.89) ~ 3.44 KB per interaction
    c = {"positives": 1000, "negatives_drawn": 2700, "updates": 890}
    assert abs(B.algorithmic_bytes(c, 64) / 1000 - 3440) < 40


def test_clock_sampler_parses_nvidia_smi_rows(tmp_path):
    s = B.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0,
                            "kill": lambda self: None})()
    s.path = str(tmp_path / "clocks.csv")
    now = time.time()
    stamp = lambda t: time.strftime("%Y/%m/%d %H:%M:%S", time.localtime(t)) + ".%03d" % int((t % 1) * 1000)
    rows = [
        "%s, 0, 1965, 1965, 600.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active" % stamp(now + 0.2),
        "%s, 0, 1950, 1965, 610.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active" % stamp(now + 0.4),
        "%s, 0, 300, 1965, 90.0, 0x0000000000000000, Not Active, Active, Not Active, Not Active" % stamp(now + 5.0),
    ]
    open(s.path, "w").write("\n".join(rows) + "\n")
    s.t0, s.t1 = now, now + 1.0   # the third row lies outside the timed region
    out = s.stop()
    assert out["samples"] == 2 and out["sm_max_mhz"] == 1965.0
    assert out["sm_mhz"] == (1965.0 + 1950.0) / 2
    assert out["reasons"] == ["sw_power_cap"]


def test_ncu_dram_csv_parser():
    import bench
    text = ('"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size",'
            '"Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
            '"0","1","python","h","fast_slot_kernel","1","7","(256, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics",'
            '"dram__bytes_read.sum","Gbyte","6.12"\n'
            '"0","1","python","h","fast_slot_kernel","1","7","(256, 1, 1)","(444, 1, 1)","0","10.0","Command line profiler metrics",'
            '"dram__bytes_write.sum","Mbyte","3,499.5"\n')
    total, seen = bench.parse_ncu_dram_csv(text)
    assert seen == 2 and abs(total - (6.12e9 + 3499.5e6)) < 1.0
    assert bench.parse_ncu_dram_csv("==PROF== nothing\n") == (0.0, 0)
