package muse

/*
#include "muse_b200.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Sharded comparison groups: one process per GPU of a box (MUSE_DEVICE = local rank), each holding a contiguous
// block of the group's series; the reference is given to every process.  go-muse has no such mode -- its Batch.Run
// walks one in-memory Group (muse_batch.go:99-130) -- so this file is an addition, not a mirror: the scores,
// the group max before the filter (muse_batch.go:87-89), the filter (results.go:46-52) and the top-N
// (results.go:54-87) of the WHOLE group come out of every process identical to a single-process Run.
//
// A step is one call, Batch.RunSharded: the kernels that produce a shard's records store them into every process's
// receive buffer over NVLink peer memory and the merge runs on the device (muse_batch_run_exchange_ex); there is no
// collective library underneath.  Setup needs one out-of-band exchange of 64 bytes per process (Handle /
// OpenPeers), e.g. over the launcher's rendezvous.

// ExchangeHandleBytes is the size of the handle every process publishes to its peers.
const ExchangeHandleBytes = 64

// Exchange is this process's end of the peer-memory exchange.
type Exchange struct {
	x     *C.muse_exchange
	rank  int
	world int
}

// NewExchange creates the receive buffers for `capacity` records per peer and step: at least TopN for runs without
// groupByLabels, at least the number of label groups a shard can hold for grouped runs.
func NewExchange(rank, world, capacity int) (*Exchange, error) {
	c, err := deviceContext()
	if err != nil {
		return nil, err
	}
	if world < 1 || rank < 0 || rank >= world {
		return nil, fmt.Errorf("muse_b200: rank %d of %d", rank, world)
	}
	e := &Exchange{rank: rank, world: world}
	if rc := C.muse_exchange_create(c, C.int32_t(rank), C.int32_t(world), C.int64_t(capacity), &e.x); rc != C.MUSE_OK {
		return nil, lastError(rc)
	}
	runtime.SetFinalizer(e, func(e *Exchange) { e.Close() })
	return e, nil
}

// Handle returns the 64 bytes the other processes need to map this process's receive buffers.
func (e *Exchange) Handle() ([]byte, error) {
	h := make([]byte, ExchangeHandleBytes)
	if rc := C.muse_exchange_ipc_handle(e.x, unsafe.Pointer(&h[0])); rc != C.MUSE_OK {
		return nil, lastError(rc)
	}
	return h, nil
}

// OpenPeers maps every process's buffers; handles[r] is what rank r's Handle returned.
func (e *Exchange) OpenPeers(handles [][]byte) error {
	if len(handles) != e.world {
		return fmt.Errorf("muse_b200: %d handles for %d ranks", len(handles), e.world)
	}
	all := make([]byte, 0, e.world*ExchangeHandleBytes)
	for r, h := range handles {
		if len(h) != ExchangeHandleBytes {
			return fmt.Errorf("muse_b200: handle of rank %d has %d bytes", r, len(h))
		}
		all = append(all, h...)
	}
	if rc := C.muse_exchange_open_peers(e.x, unsafe.Pointer(&all[0])); rc != C.MUSE_OK {
		return lastError(rc)
	}
	return nil
}

// Close releases the buffers (peers must have stopped stepping).
func (e *Exchange) Close() {
	if e.x != nil {
		C.muse_exchange_destroy(e.x)
		e.x = nil
	}
}

// SetShard declares this Group to be the block of a sharded group that starts at series index `first` of the whole,
// and fixes the label dictionary: for every label key the complete list of its values, IDENTICAL on every
// process, so that the int32 ids the device groups by mean the same labels everywhere.  Call before the first
// NewBatch on the group.
func (g *Group) SetShard(first int64, dictionary map[string][]string) {
	g.shardFirst = first
	g.fixedDict = dictionary
	if g.store != nil { // re-encode with the fixed ids
		C.muse_group_destroy(g.store)
		g.store = nil
		g.uploaded = 0
	}
}

// ShardScore is one score of the whole sharded group.  Index counts series over all shards in rank order; Labels is
// set when that series lives in this process's shard (the other processes hold the rest).
type ShardScore struct {
	Index        int64
	Lag          int
	PercentScore float64
	Labels       *Labels
}

// RunSharded is Batch.Run over the whole sharded group: every process calls it the same number of times with the same
// groupByLabels and Results settings, and every process gets the same scores back, best first (results.go:54-87).
func (b *Batch) RunSharded(e *Exchange, groupByLabels []string) ([]ShardScore, error) {
	comp := b.Comparison
	if len(comp.series) == 0 {
		return nil, errNoSeries
	}
	if err := b.bind(); err != nil {
		return nil, err
	}
	r := b.Results
	if r.TopN < 1 {
		return nil, nil
	}
	cols := comp.keyCols(groupByLabels)
	scores := make([]float64, r.TopN)
	lags := make([]int64, r.TopN)
	idx := make([]int64, r.TopN)
	var nOut C.int64_t
	var colPtr *C.int32_t
	if len(cols) > 0 {
		colPtr = (*C.int32_t)(unsafe.Pointer(&cols[0]))
	}
	rc := C.muse_batch_run_exchange_ex(b.batch, e.x, colPtr, C.int32_t(len(cols)), C.int64_t(r.MaxLag), C.int64_t(r.TopN),
		C.double(r.Threshold), C.int32_t(r.SignFilter), C.MUSE_MODE_AUTO,
		(*C.double)(unsafe.Pointer(&scores[0])), (*C.int64_t)(unsafe.Pointer(&lags[0])),
		(*C.int64_t)(unsafe.Pointer(&idx[0])), &nOut)
	if rc != C.MUSE_OK {
		return nil, lastError(rc)
	}
	out := make([]ShardScore, int(nOut))
	for i := range out {
		out[i] = ShardScore{Index: idx[i], Lag: int(lags[i]), PercentScore: scores[i]}
		if local := idx[i] - comp.shardFirst; local >= 0 && local < int64(len(comp.series)) {
			out[i].Labels = comp.series[local].Labels()
		}
	}
	return out, nil
}
