#!/bin/bash
# in-graph marginal cost per kernel class: step time with one class left out (diagnostic)
for s in none ln attn conv decode ff qkv out pw sub mel "ln,attn,conv,ff,qkv,out,pw" ; do
  echo -n "skip=$s : "; NSB_SKIP=$s python tools/ncu_step.py 6 2>&1 | tail -1
done
