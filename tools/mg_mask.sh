N=${1:-2}
for v in "PAGNERF_SYMM_P2P=1 PAGNERF_SYNC_MASK=2 PAGNERF_GRAD_TRANSPORT=symm" "PAGNERF_SYMM_P2P=1 PAGNERF_SYNC_MASK=7 PAGNERF_GRAD_TRANSPORT=symm"; do
echo "== $v"
env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/mg_mask.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); k=j['kernels_ms_per_step']; print(j['ms_per_step'], j.get('allreduce')['exposed_ms'], k.get('pag_allreduce_symm'), k.get('pag_symm_barrier'))
" || tail -5 gpurun_out/mg_mask.err
done
