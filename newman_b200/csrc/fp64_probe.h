// FP64 issue-rate probes (fp64_probe.cu): measurement code behind nm_fp64_peak, kept out of the hot-path TU.
#pragma once
int nm_probe_run(void* cuda_stream, int sm_count, int kind, int iters, double* sink, double* inst_per_s, double* ms_out);
