/* pnp-b200 — forwarding header for `#include "fasp4ns_functs.h"`: the prototypes live in fasp_compat.h */
#ifndef PNPB200_COMPAT_FASP4NS_FUNCTS_H
#define PNPB200_COMPAT_FASP4NS_FUNCTS_H
#include "../fasp_compat.h"
#endif
