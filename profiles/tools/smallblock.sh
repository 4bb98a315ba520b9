#!/bin/bash
# The reference's call granularity from a native (C++) caller: one stream, 65536-sample blocks, pageable vectors.
#   blocks   Downsampler::process + UDPSinkFEC::write (the mirror classes, synchronous per block)
#   blocksq  sdrd_rx_submit + sdrd_rx_collect (queued), min chain 0 (an idle device starts at once) and 16 blocks
cd "$(dirname "$0")/../.." || exit 1
(cd tests && python -c "import host_cases; print(host_cases.build('gpu'))")
for d in 4 1 6; do ./tests/host/host_pipeline_gpu blocks 2000 $d 16 65536 50; done
for mc in 0 4 16; do ./tests/host/host_pipeline_gpu blocksq 4000 4 16 65536 $mc 64; done
./tests/host/host_pipeline_gpu blocksq 4000 6 32 65536 16 64
