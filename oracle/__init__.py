"""CPU oracle for the FollowYourClick denoising hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain fp32 PyTorch-CPU restatement of the
reference algorithm.  It exists so that the CUDA engine
in ``followyourclick_b200`` can be checked on a GPU box where ``/root/reference`` does not
exist.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does.

Pinning: the reference ships no tests/golden vectors for this path,
so the oracle is pinned against *outputs of the reference itself*, produced in the build
container by ``tests/golden/make_golden.py`` (imports the unmodified reference from
/root/reference with the three import shims of SURVEY App. C, loads the same synthetic
state dict, and stores the reference outputs as fixtures under ``tests/golden/``).
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures on every run.
"""
