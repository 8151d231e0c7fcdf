#include "cpu.h"
#include <stdio.h>
#include <string.h>
#include <strings.h>

static char model_name[256] = "";
static ucc_cpu_vendor_t vendor = UCC_CPU_VENDOR_LAST;

static void detect(void)
{
    FILE *f = fopen("/proc/cpuinfo", "r"); char line[512];
    vendor = UCC_CPU_VENDOR_UNKNOWN;
    if (!f) return;
    while (fgets(line, sizeof(line), f)) {
        if (!strncmp(line, "vendor_id", 9)) {
            if (strstr(line, "GenuineIntel")) vendor = UCC_CPU_VENDOR_INTEL;
            else if (strstr(line, "AuthenticAMD")) vendor = UCC_CPU_VENDOR_AMD;
        } else if (!strncmp(line, "CPU implementer", 15)) {
            vendor = strstr(line, "0x4e") ? UCC_CPU_VENDOR_NVIDIA : UCC_CPU_VENDOR_ARM;
        } else if (!strncmp(line, "model name", 10) && !model_name[0]) {
            char *c = strchr(line, ':');
            if (c) { snprintf(model_name, sizeof(model_name), "%s", c + 2); model_name[strcspn(model_name, "\n")] = 0; }
        }
    }
    fclose(f);
}
ucc_cpu_vendor_t ucc_arch_get_cpu_vendor(void) { if (vendor == UCC_CPU_VENDOR_LAST) detect(); return vendor; }
const char *ucc_cpu_vendor_string(ucc_cpu_vendor_t v)
{ static const char *n[] = {"unknown", "intel", "amd", "arm", "nvidia"}; return v < UCC_CPU_VENDOR_LAST ? n[v] : "unknown"; }
const char *ucc_arch_get_cpu_model_string(void)
{
    static const struct { const char *pat, *name; } map[] = {
        {"EPYC 7", "rome"}, {"EPYC 9", "genoa"}, {"Platinum 81", "skylake"}, {"Platinum 82", "cascadelake"},
        {"Platinum 83", "icelake"}, {"Platinum 84", "sapphirerapids"}, {"Platinum 85", "emeraldrapids"},
        {"Grace", "grace"}, {"Neoverse", "neoverse"}, {NULL, NULL}};
    if (vendor == UCC_CPU_VENDOR_LAST) detect();
    for (int i = 0; map[i].pat; i++) if (strstr(model_name, map[i].pat)) return map[i].name;
    return "unknown";
}
