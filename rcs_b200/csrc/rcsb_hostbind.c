/*
 * rcsb_bind_host_to_device: NUMA placement of the host side of the RCSB_HOST_PTRS path.
 *
 * With one rank per GPU on a two-socket box, the copies between a rank's pinned buffers and its
 * GPU cross the inter-socket link unless the buffers live on the socket the GPU hangs off. The
 * topology is read from sysfs (PCI device -> numa_node -> cpulist); affinity and the preferred
 * memory node are set with the plain Linux system calls (no libnuma dependency).
 */
#include "rcsb.h"
#include "rcsb_internal.h"

#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

/* implemented in rcsb_api.cu (CUDA runtime): "0000:1b:00.0" of a device, lower case */
int rcsb_device_pci_bus_id(int device, char* busId, int len);

#define RCSB_MPOL_PREFERRED 1

static int readFirstLine(const char* path, char* buf, size_t len)
{
  FILE* f = fopen(path, "r");
  if (!f)
  {
    return -1;
  }
  if (!fgets(buf, (int) len, f))
  {
    fclose(f);
    return -1;
  }
  fclose(f);
  buf[strcspn(buf, "\r\n")] = 0;
  return 0;
}

/* "0-31,64-95" -> cpu set */
static int parseCpuList(const char* s, cpu_set_t* set)
{
  int n = 0;
  CPU_ZERO(set);
  while (*s)
  {
    char* end;
    long a = strtol(s, &end, 10);
    if (end == s)
    {
      break;
    }
    long b = a;
    if (*end == '-')
    {
      s = end + 1;
      b = strtol(s, &end, 10);
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
    {
      CPU_SET((int) c, set);
      ++n;
    }
    s = (*end == ',') ? end + 1 : end;
    if (*end != ',' && *end != 0)
    {
      break;
    }
  }
  return n;
}

int rcsb_bind_host_to_device(int device, char* info, size_t infoBytes)
{
  char bus[64], path[256], line[4096];
  char msg[512];
  msg[0] = 0;
  if (info && infoBytes > 0)
  {
    info[0] = 0;
  }
  if (rcsb_device_pci_bus_id(device, bus, (int) sizeof(bus)) != 0)
  {
    rcsb_set_error("rcsb_bind_host_to_device: unknown device %d", device);
    return RCSB_EINVAL;
  }

  int node = -1;
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  if (readFirstLine(path, line, sizeof(line)) == 0)
  {
    node = atoi(line);
  }
  if (node < 0)
  {
    snprintf(msg, sizeof(msg), "gpu %d (%s): no NUMA node in sysfs, nothing bound", device, bus);
  }
  else
  {
    cpu_set_t nodeCpus, allowed, both;
    int nBoth = 0, cpuBound = 0, memBound = 0;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    if (readFirstLine(path, line, sizeof(line)) == 0 && parseCpuList(line, &nodeCpus) > 0 &&
        sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
    {
      CPU_AND(&both, &nodeCpus, &allowed);
      nBoth = CPU_COUNT(&both);
      if (nBoth > 0 && nBoth < CPU_COUNT(&allowed))
      {
        cpuBound = sched_setaffinity(0, sizeof(both), &both) == 0;
      }
    }
    /* preferred (not strict) policy: falls back to other nodes when the cpuset forbids this one */
    if (node < 1024)
    {
      unsigned long mask[16];
      memset(mask, 0, sizeof(mask));
      mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
      memBound = syscall(SYS_set_mempolicy, RCSB_MPOL_PREFERRED, mask, 8 * sizeof(mask) + 1) == 0;
    }
    snprintf(msg, sizeof(msg),
             "gpu %d (%s) on NUMA node %d: %d of its CPUs allowed, thread affinity %s, preferred memory node %s",
             device, bus, node, nBoth, cpuBound ? "set" : "unchanged", memBound ? "set" : "unchanged");
  }
  if (info && infoBytes > 0)
  {
    snprintf(info, infoBytes, "%s", msg);
  }
  return RCSB_OK;
}
