#!/bin/bash
O=gpurun_out/t; mkdir -p $O
timeout 600 python tools/r02/dbg_fx.py > $O/dbg.log 2>&1; tail -30 $O/dbg.log
