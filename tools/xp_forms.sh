for form in 0 1; do for hs in 0 1 2 3; do
rm -f gpurun_out/parity_achieved.jsonl
DPGP_BOUND_FORM=$form DPGP_HSYM=$hs python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(test_objective_and_gradients_vs_reference and c1) or shape13 or shape0" > /dev/null 2>&1
echo "form=$form hsym=$hs" >> gpurun_out/xp_forms.txt
python - >> gpurun_out/xp_forms.txt <<'PY'
import json
for l in open('gpurun_out/parity_achieved.jsonl'):
    r=json.loads(l)
    w=max(r['grad_rel_err'].items(), key=lambda kv: kv[1])
    print("  %-34s kappa %.1e obj %.1e worst %s %.1e"%(r['case'],r['kappa'],r['objective_rel_err'],w[0],w[1]))
PY
done; done
cat gpurun_out/xp_forms.txt
