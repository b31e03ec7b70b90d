package muse

/*
#include "muse_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math"
	"runtime"
	"unsafe"
)

// Muse runs a z-normalized cross correlation between a reference series and the series of ONE
// group per Run call, keeping the sign of the peak (go-muse muse.go:12-92).  The scores come from
// muse_batch_score_all(signed_scores = 1): clamped to [-1, 1] as muse.go:72-76.
type Muse struct {
	refN    int
	n       int
	ref     []float64
	store   *C.muse_group
	batch   *C.muse_batch
	Results *Results
}

// New creates a Muse instance with a reference timeseries and results (muse.go:23-45).
func New(ref *Series, results *Results) (*Muse, error) {
	if ref.Length() < 1 {
		return nil, errors.New("Reference series length must be greater than zero")
	}
	c, err := deviceContext()
	if err != nil {
		return nil, err
	}
	m := &Muse{refN: ref.Length(), ref: append([]float64(nil), ref.Values()...), Results: results}
	if rc := C.muse_group_create(c, C.int64_t(m.refN), 0, 1, &m.store); rc != C.MUSE_OK {
		return nil, lastError(rc)
	}
	rc := C.muse_batch_create(c, m.store, (*C.double)(unsafe.Pointer(&m.ref[0])), C.int64_t(m.refN), &m.batch)
	if rc != C.MUSE_OK {
		C.muse_group_destroy(m.store)
		if rc == C.MUSE_ERR_STDDEV_ZERO {
			return nil, fmt.Errorf("Invalid input query, %v", "Standard deviation of zero")
		}
		return nil, lastError(rc)
	}
	m.n = int(C.muse_batch_fft_len(m.batch))
	runtime.SetFinalizer(m, func(m *Muse) {
		C.muse_batch_destroy(m.batch)
		C.muse_group_destroy(m.store)
	})
	return m, nil
}

// Run compares the comparison series against the reference and updates the results with the one
// that has the largest |score| (strictly greater wins, the first is always taken: muse.go:86).
func (m *Muse) Run(compGraphs []*Series) error {
	if len(compGraphs) == 0 {
		return nil
	}
	rows := make([]float64, 0, len(compGraphs)*m.refN)
	for _, s := range compGraphs {
		if s.Length() != m.refN {
			return fmt.Errorf("Encountered a comparison graph with differing length than the reference, %+v", s.Labels())
		}
		rows = append(rows, s.Values()...)
	}
	if rc := C.muse_group_clear(m.store); rc != C.MUSE_OK {
		return lastError(rc)
	}
	if rc := C.muse_group_append(m.store, (*C.double)(unsafe.Pointer(&rows[0])), C.int64_t(len(compGraphs)),
		C.int64_t(m.refN), nil); rc != C.MUSE_OK {
		return lastError(rc)
	}
	scores := make([]float64, len(compGraphs))
	lags := make([]int32, len(compGraphs))
	if rc := C.muse_batch_score_all(m.batch, 1, (*C.double)(unsafe.Pointer(&scores[0])),
		(*C.int32_t)(unsafe.Pointer(&lags[0]))); rc != C.MUSE_OK {
		return lastError(rc)
	}
	best := 0
	for i := 1; i < len(scores); i++ {
		if math.Abs(scores[i]) > math.Abs(scores[best]) {
			best = i
		}
	}
	m.Results.Update(Score{Labels: compGraphs[best].Labels(), Lag: int(lags[best]), PercentScore: scores[best]})
	return nil
}
