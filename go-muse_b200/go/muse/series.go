package muse

import "github.com/google/uuid"

// Series is a timeseries of values with labels (go-muse series.go:8-42).
type Series struct {
	y      []float64
	labels *Labels
}

// NewSeries creates a Series; without labels a unique "uid" label is generated.
func NewSeries(y []float64, labels *Labels) *Series {
	if labels == nil || labels.Len() == 0 {
		labels = NewLabels(LabelMap{DefaultLabel: uuid.New().String()})
	}
	return &Series{y: y, labels: labels}
}

// Length returns the length of the timeseries.
func (s *Series) Length() int { return len(s.y) }

// Values returns the series values (never modified by this package).
func (s *Series) Values() []float64 { return s.y }

// Labels returns the labels of the timeseries.
func (s *Series) Labels() *Labels { return s.labels }

// UID is the unique identifier of the series within a Group.
func (s *Series) UID() string { return s.labels.ID(s.labels.Keys()) }
