#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/tools/two_lane_experiment.py > gpurun_out/r2m_two_lane.json 2> gpurun_out/r2m_two_lane.err; echo rc=$?
cat gpurun_out/r2m_two_lane.json; tail -5 gpurun_out/r2m_two_lane.err
