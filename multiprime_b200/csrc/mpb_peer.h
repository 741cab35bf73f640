// mpb_peer.h — internal interface of the peer-memory all-reduce (mpb_peer.cu), shared with the device walk
#pragma once
#include <stdint.h>

#include "mpb200.h"
#include "mpb_host.h"

#define MPB_ERR_PEER_CAP 16      // walk error flags (mpb_walk_dev.cu): vector longer than the receive slots
#define MPB_ERR_PEER_TIMEOUT 32  // a peer's signal did not arrive

// Enqueue the all-reduce of data_d[0 .. n_items_d[0] * mult) on the context's stream (no synchronisation); errors are
// OR-ed into err_d.
int mpb_peer_allreduce_launch(mpb_peer* p, unsigned long long* data_d, const int* n_items_d, int mult, int* err_d);
