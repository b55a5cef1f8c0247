// Keeps the reference's module path so that main/ imports resolve unchanged
// (reference go.mod:1): drop this directory's pkg/fluid over the reference's, or point a
// `replace github.com/TheFellow/fluid => <repo>/go` directive at it.
module github.com/TheFellow/fluid

go 1.24.2
