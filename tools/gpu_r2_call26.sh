#!/bin/bash
# three steps in flight: the API / production-path / serve tests after the test updates, then the whole suite
O=gpurun_out/c26; mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
