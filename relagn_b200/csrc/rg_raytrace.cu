// rg_raytrace.cu -- device Kerr ray tracer / transfer-table generator (placeholder until the
// tracer lands: tables can be uploaded with rg_tables_upload meanwhile).
#include "rg_kernels.h"

int rg_launch_raytrace(const double *, int, const double *, int, int, int, double, float *, long *, cudaStream_t)
{
    return rg_set_error(RG_E_UNSUPPORTED, "rg_tables_generate: device ray tracer not built into this library yet");
}
