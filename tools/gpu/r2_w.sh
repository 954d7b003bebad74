#!/bin/bash
# round 2, call W: the default bench line of the final tree
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2w; mkdir -p $O
timeout 140 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
