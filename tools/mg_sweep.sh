N=${1:-2}
for v in ${SWEEP:-"PAGNERF_SYMM_SPLIT_LEVEL=18" "PAGNERF_SYMM_SPLIT_LEVEL=21"}; do
echo "== $v"
env $v PAGNERF_GRAD_TRANSPORT=symm timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/mg.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); k=j['kernels_ms_per_step']; print(j['ms_per_step'], j['value'], j.get('allreduce')['exposed_ms'], k.get('pag_allreduce_symm'), k.get('pag_symm_barrier'), k.get('pag_permuto_bwd_img16_dyn'))
" || tail -5 gpurun_out/mg.err
done
