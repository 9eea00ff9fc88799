mkdir -p gpurun_out
KNNSVC_NVCC_EXTRA=-DKNNSVC_K5_PROFILE python -m knn_svc_b200.build --force > /dev/null 2>&1 || echo "build failed"
echo "== cluster" ; timeout 120 python tools/k5_profile.py 2>&1 | tail -12
echo "== one CTA" ; KNNSVC_OPTIONS=concat_cluster=0 timeout 120 python tools/k5_profile.py 2>&1 | tail -12
