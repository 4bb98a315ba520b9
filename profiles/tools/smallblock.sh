cd tests && python -c "
import host_cases; print(host_cases.build('gpu'))" && cd ..
for d in 4 1 6; do ./tests/host/host_pipeline_gpu blocks 2000 $d 16 65536 50; done
./tests/host/host_pipeline_gpu blocks 2000 4 16 16384 50
./tests/host/host_pipeline_gpu blocks 500 4 16 1048576 20
