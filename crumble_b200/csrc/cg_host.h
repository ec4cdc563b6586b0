/* internal declarations shared by cg_host.cpp, cg_device.cu and the test-only emulator */
#ifndef CG_HOST_H
#define CG_HOST_H
#include "cg_pipeline.h"
int  cg_params_check(const cg_params *p, const char **why);
int  cg_params_generic(const cg_params *p);
void cg_bed_prefix_max(const cg_bed_reg *bed, int n, int64_t *pm);
void cg_devparams_from(CgDevParams *d, const cg_params *p);
void cg_tables_init(CgTables *T, const cg_params *p);
extern "C" int cg_enable_pinned(void);   /* installs the pinned-memory hooks when a device exists */
extern void *(*cg_pinned_alloc_hook)(size_t);
extern void (*cg_pinned_free_hook)(void *);
#endif
