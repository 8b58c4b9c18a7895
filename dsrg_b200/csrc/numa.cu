// NUMA placement of the host side of one-process-per-GPU jobs (Linux, no libnuma).
//
// The host-buffer entry points stream ~1.7 GB per 64-image step through host DRAM (pinned blobs in, seeds out,
// wire.cu).  On the 8-GPU boxes GPUs 0-3 hang off NUMA node 0 and 4-7 off node 1; a rank whose pinned buffers and
// packing threads live on the other socket pays the inter-socket link for every byte (round 1: end-to-end scaling
// 0.43 at 8 GPUs while the device-resident path scaled 0.997).  When several ranks share the host
// (LOCAL_WORLD_SIZE > 1, or DSRG_B200_NUMA=1) the engine therefore binds the calling thread -- and with it the
// OpenMP team it creates later -- to the CPUs of its GPU's node and allocates its pinned memory there.
// DSRG_B200_NUMA=0 turns all of it off.  The reference's counterpart is the process fan-out of
// multiprocessing.Pool (pylayers/pylayers/pylayers.py:292), which leaves placement to the OS.
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "common.cuh"

namespace dsrg {

static int read_int_file(const char *path, int fallback) {
    FILE *f = fopen(path, "r");
    if (!f) return fallback;
    int v = fallback;
    if (fscanf(f, "%d", &v) != 1) v = fallback;
    fclose(f);
    return v;
}

bool numa_wanted() {
    if (const char *ev = getenv("DSRG_B200_NUMA")) return atoi(ev) != 0;
    if (const char *ev = getenv("LOCAL_WORLD_SIZE")) return atoi(ev) > 1;
    return false;
}

// NUMA node the GPU's PCIe root belongs to (-1: unknown / single node)
int numa_node_of_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *p = bus; *p; p++)
        if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');  // sysfs uses lower-case hex
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    return read_int_file(path, -1);
}

// parse a sysfs cpulist ("0-31,64-95") into a cpu_set_t; returns the number of CPUs
static int parse_cpulist(const char *path, cpu_set_t *set) {
    CPU_ZERO(set);
    FILE *f = fopen(path, "r");
    if (!f) return 0;
    char buf[4096] = {0};
    const size_t n = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[n] = 0;
    int count = 0;
    for (char *p = buf; *p;) {
        char *end;
        long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        if (*end == '-') b = strtol(end + 1, &end, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, set);
            count++;
        }
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',' ) break;
    }
    return count;
}

// restrict the calling thread (and the threads it creates from now on) to the node's CPUs that it may already use
bool numa_bind_thread(int node) {
    if (node < 0) return false;
    char path[128];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    cpu_set_t want, have, both;
    if (parse_cpulist(path, &want) == 0) return false;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return false;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return false;
    return sched_setaffinity(0, sizeof(both), &both) == 0;
}

// memory policy of the calling thread: prefer `node` (node < 0: back to the default policy)
bool numa_prefer_memory(int node) {
#ifdef SYS_set_mempolicy
    const int MPOL_DEFAULT_ = 0, MPOL_PREFERRED_ = 1;
    if (node < 0) return syscall(SYS_set_mempolicy, MPOL_DEFAULT_, nullptr, 0) == 0;
    unsigned long mask[16] = {0};
    if (node >= (int)(sizeof(mask) * 8)) return false;
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    return syscall(SYS_set_mempolicy, MPOL_PREFERRED_, mask, sizeof(mask) * 8) == 0;
#else
    (void)node;
    return false;
#endif
}

// pinned allocation on the node of `device` (plain cudaHostAlloc when placement is off or unknown)
cudaError_t numa_host_alloc(void **p, size_t bytes, int device) {
    const int node = (numa_wanted() && device >= 0) ? numa_node_of_device(device) : -1;
    const bool pol = node >= 0 && numa_prefer_memory(node);
    const cudaError_t err = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (pol) numa_prefer_memory(-1);
    return err;
}

}  // namespace dsrg
