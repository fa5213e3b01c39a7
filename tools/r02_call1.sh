# round-2 regression on one GPU: every -m gpu test, smoke, the default bench (both arms)
cd /root/repo
tag=${1:-r02_a}
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
( time python __graft_entry__.py smoke ) > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1; tail -4 gpurun_out/${tag}_bench.log | cut -c1-3000
