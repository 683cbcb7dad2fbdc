#!/bin/bash
for k in 13 10; do timeout 60 build/climb_exp_128 $k base | tail -2; done
