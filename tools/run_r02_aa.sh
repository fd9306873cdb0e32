#!/bin/bash
timeout 300 python tools/careful_probe.py 2>&1 | tail -6
