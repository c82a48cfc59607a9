# call C: spread TMA zero fill; ncu --set full of pcl_compact and tac_contact
mkdir -p gpurun_out
bash tools/gpu_variants2.sh ""
bash tools/gpu_ncu1.sh pcl_compact_kernel c_compact
bash tools/gpu_ncu1.sh tac_contact c_contact
