#!/bin/bash
# Scratch copy of the tree under build/work, so that /root/repo (which gpurun snapshots at an unpredictable moment,
# when a GPU slot frees up) is only ever updated in one quick, consistent step:
#   tools/work.sh out    copy the repo into build/work (start of an editing session)
#   tools/work.sh back   build the library in build/work if stale, then copy sources + the built .so back
# (files deleted in build/work are not deleted in the repo: remove those by hand)
ROOT=/root/repo
W=$ROOT/build/work
EX="--exclude=./.git --exclude=./build --exclude=./gpurun_out --exclude=__pycache__ --exclude=./.pytest_cache"
case "$1" in
  out)  rm -rf $W && mkdir -p $W && tar -C $ROOT $EX -cf - . | tar -C $W -xf - ;;
  back) (cd $W && python -m sgmcmcjax_b200._build > /dev/null 2>$W/../work_build.err) || { echo "build failed:"; grep -E "error" -A3 $W/../work_build.err | head -30; exit 1; }
        tar -C $W $EX -cf - . | tar -C $ROOT -xf - && touch $ROOT/sgmcmcjax_b200/libsgmcmc_b200.so && echo synced ;;
  *) echo "usage: tools/work.sh out|back"; exit 2 ;;
esac
