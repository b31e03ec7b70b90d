package muse

import "sort"

// DefaultLabel is the label name if a series is specified without any labels (go-muse labels.go:7).
const DefaultLabel = "uid"

// LabelMap is a map of label keys to values.
type LabelMap map[string]string

// Labels is a map of label names to label values with its sorted keys cached.
type Labels struct {
	labels LabelMap
	keys   []string
}

// NewLabels creates a Label storing the map and sorted keys (go-muse labels.go:20-30).
func NewLabels(labels LabelMap) *Labels {
	l := &Labels{labels: labels, keys: make([]string, 0, len(labels))}
	for k := range labels {
		l.keys = append(l.keys, k)
	}
	sort.Strings(l.keys)
	return l
}

// Len returns the number of labels.
func (l *Labels) Len() int { return len(l.labels) }

// Keys returns the sorted keys of the labels.
func (l *Labels) Keys() []string { return l.keys }

// Get returns the value of a specified key and whether it is present.
func (l *Labels) Get(key string) (string, bool) {
	v, ok := l.labels[key]
	return v, ok
}

// ID builds "key1:val1,key2:val2" over the sorted requested keys, skipping absent keys; like
// go-muse (labels.go:54-73) it sorts the caller's slice in place.
func (l Labels) ID(labels []string) string {
	if len(labels) == 0 {
		labels = l.Keys()
	} else {
		sort.Strings(labels)
	}
	out := ""
	for _, k := range labels {
		if v, ok := l.Get(k); ok {
			out += k + ":" + v + ","
		}
	}
	if len(out) > 0 {
		return out[:len(out)-1]
	}
	return out
}
