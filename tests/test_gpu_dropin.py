"""The drop-in, run: oracle/_ref/pf_ref_gpu is the REFERENCE's own JobTree::run_master, records and NormalSampler<2>
(compiled in place from the reference sources, oracle/Makefile `ref-gpu`) over the product's frontier communicator and
`pf_eval_frontier` of libpfetch.so -- the binding INTEGRATION.md section 2a prints, as a binary.  Its chain must be the
chain the reference's own MasterTester harness produced on the CPU (tests/golden/chain_normal_d10000.tsv): identical
accept / reject decisions and theta bytes, metric sums within 1e-10."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "pf_ref_gpu")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("frontier", [16, 1, 64])
def test_reference_tree_over_the_gpu_engine_reproduces_the_reference_chain(orc, frontier):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/pf_ref_gpu was not built (needs /root/reference at build time)")
    p = subprocess.run([EXE, "normal", "-W", str(frontier), "-d", "10000", "-l", "300"], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    got = orc.parse_chain(p.stdout)
    gold = orc.parse_chain(open(os.path.join(GOLD, "chain_normal_d10000.tsv")).read())
    assert len(got) == len(gold) == 301
    for i, (a, g) in enumerate(zip(got, gold)):
        assert (a[0], a[1], a[3]) == (g[0], g[1], g[3]), i          # depth, accept, theta bytes
        assert abs(a[2] - g[2]) <= 1e-10 * abs(g[2]), i
    tail = [l for l in p.stdout.splitlines() if l.startswith("# levels")][0].split()
    stats = dict(zip(tail[1::2], tail[2::2]))
    # one kernel pass per scheduling burst: a launch, or a command served by the resident kernel (launched once)
    passes = int(stats["kernel_launches"]) - int(stats["resident_launches"]) + int(stats["resident_calls"])
    assert passes == int(stats["engine_calls"]) >= 1
    assert int(stats["max_batch"]) <= frontier
