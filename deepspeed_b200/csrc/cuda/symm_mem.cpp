// placeholder
extern "C" __attribute__((visibility("default"))) int dsb_symm_mem_version() { return 0; }
