#!/bin/bash
export BMC_B200_LIB=$PWD/bmcnet_esr_b200/libbmc_b200_measure.so
BMC_FRONT_PROF=1 BMC_NO_GRAPH=1 timeout 300 python tools/prof_step.py plain_nfs 95 4 2>&1 | grep frontprof | head -4
