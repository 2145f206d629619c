"""CPU oracle for the recurrent off-policy MARL learner hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `off-policy_b200/` (the product) imports this
package.  The only legitimate users are `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` (as the checker / the CPU arm,
never as the thing shipped).

What it is: a restatement, in NumPy (buffer, RNG, segment trees) and CPU PyTorch (the
learner maths — the reference's arithmetic *is* PyTorch), of the
reference functions listed in SURVEY.md §8(a).  Every function cites the reference
file:line it follows.

Pinning status: the reference ships NO tests, golden vectors or known-answer fixtures
(SURVEY.md §4), so there is nothing of the reference's own to pin against.  Instead the
oracle is pinned against *outputs of the unmodified reference itself, run in the build
container* (`tests/golden/make_goldens.py` imports /root/reference and dumps
`tests/golden/*.npz`; `tests/test_oracle_vs_golden.py` replays them through this
package).  NumPy's legacy MT19937 stream is additionally pinned against the installed
numpy (the third-party dependency the reference calls, `numpy==1.18.5` pinned in its
requirements.txt:77; the legacy stream is frozen by NumPy policy).
"""
