"""Worker process of tests.util.scipy_reference: reads (workload spec, jobs) as a pickle on stdin, runs the reference solves
(oracle.slsqp_solve == srv.py:363-364) and writes the results as a pickle on stdout.  Test infrastructure."""
import pickle
import sys

if __name__ == "__main__":
    from tests import util
    spec, jobs = pickle.loads(sys.stdin.buffer.read())
    util._pw_init(*spec)
    out = [util._pw_solve(j) for j in jobs]
    sys.stdout.buffer.write(pickle.dumps(out))
