#!/bin/bash
# round 2, session 10: where vd_update's single-CTA phase spends its time (clock stamps), the write probe for the
# unexplained DRAM reads of the sampling kernel, L2 / local-memory detail of the sampling kernel
tag=r02s10
mkdir -p gpurun_out
python profiles/vd_clocks.py > gpurun_out/${tag}_vd_clocks.txt 2>&1
cat gpurun_out/${tag}_vd_clocks.txt
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/write_probe profiles/write_probe.cu
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,lts__t_sectors_srcunit_tex_aperture_device_op_write.sum,lts__t_sectors_aperture_device_lookup_miss.sum,dram__sectors_read.sum,dram__sectors_write.sum
timeout 300 ncu --metrics $M --clock-control none -s 5 -c 5 --csv --log-file gpurun_out/${tag}_write_probe.csv /tmp/write_probe > gpurun_out/${tag}_write_probe.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"vd_sample_eval|vd_wsum" -s 2 -c 2 --csv --log-file gpurun_out/${tag}_vd_sample_l2.csv \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_vd_sample_l2.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:"vd_sample_eval" -s 2 -c 1 --csv --log-file gpurun_out/${tag}_vd_sample_l2_nocachectl.csv \
   python profiles/prof_cfg.py vd >> gpurun_out/${tag}_vd_sample_l2.log 2>&1
du -sh gpurun_out
