// Package muse is a drop-in facade for github.com/aouyang1/go-muse's Batch.Run path that runs
// on NVIDIA B200 GPUs through libmuse_b200.so (include/muse_b200.h).
//
// The exported API (NewLabels, NewSeries, NewGroup, Group.Add, NewResults, Results.Update,
// Results.Fetch, NewBatch, Batch.Run, Score, SignFilter_*) keeps go-muse's names, argument
// meaning and error behaviour.  Differences, all documented in DESIGN.md section 2:
//   - series values are copied to the device at NewBatch/Run time and never modified
//     (go-muse z-normalises the caller's slices in place);
//   - ties resolve deterministically (highest score, then first-added series);
//   - Batch.Concurrency is accepted and ignored: the device schedules the work.
//
// NOTE: this image has no Go toolchain; the package is written against the C ABI but has not
// been compiled here.  tests/ exercise the identical logic through the Python mirror
// (go-muse_b200/muse_b200.py).
package muse
