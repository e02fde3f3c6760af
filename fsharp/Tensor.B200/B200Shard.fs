namespace Tensor.B200

// Leading-axis sharding of Tensor<'T> over the GPUs of one box (SURVEY.md §8e) — the host-side half of the
// dn_shard_* entry points (include/dn_tensor.h "Multi-GPU"). New: the reference drives one device
// (Tensor/Tensor/Cuda/CudaUtils.fs:42-46). One .NET process drives every rank (nlocal = world): a rank is a CUDA
// device, a slab of dim 0 of every operand, and a stream; per-rank work is issued rank by rank from one thread (every
// call is asynchronous) or from one thread per rank. NOT COMPILED HERE (no .NET toolchain in this image); the
// Python mirror of the same calls is deepnet_b200/shard.py, exercised by tests/test_shard_gpu.py.

open System
open Tensor
open Tensor.Backend

/// A group of `devices.Length` ranks, rank r on CUDA device devices.[r].
type ShardGroup (devices: int[], heapBytes: int64) =
    let world = devices.Length
    let mutable group = 0n
    do
        Native.check (Native.dn_shard_group_create (world, world, Array.init world id, devices, heapBytes, &group))
        Native.check (Native.dn_shard_group_connect (group, null))

    member this.World = world
    member this.Handle = group

    /// Rows [fst, fst + snd) of `nrows` owned by `rank` (the remainder goes to the first ranks).
    member this.Slab (nrows: int64, rank: int) =
        let mutable b, c = 0L, 0L
        Native.check (Native.dn_shard_slab (nrows, rank, world, &b, &c))
        b, c

    /// Device memory for a full (replicated) result in rank's symmetric heap.
    member this.HeapAlloc (rank: int, nbytes: int64) =
        let mutable p = 0n
        Native.check (Native.dn_shard_heap_alloc (group, rank, nbytes, &p))
        p

    /// One thread drives all ranks: each collective is issued on every rank inside a group bracket
    /// (dn_shard_group_start / dn_shard_group_end, cf. ncclGroupStart / ncclGroupEnd) — the kernels launch at once,
    /// the waits follow when every rank has issued.
    member private this.Bracket (f: int -> unit) =
        Native.check (Native.dn_shard_group_start group)
        try
            for r in 0 .. world - 1 do f r
        finally
            Native.check (Native.dn_shard_group_end group)

    member this.Sync () = for r in 0 .. world - 1 do Native.check (Native.dn_shard_sync (group, r))
    member this.Barrier () = this.Bracket (fun r -> Native.check (Native.dn_shard_barrier (group, r)))

    /// `op` LastAxis of every rank's slab into the full result on every rank: one kernel launch per rank, the outputs
    /// travel as peer stores over NVLink. tFull.[r] / aLocal.[r] are rank r's descriptors (full result in its heap,
    /// slab with the reduced axis last); rowBegin.[r] is the slab's first row.
    member this.ReduceLastAxis (op: int, tFull: DnTensor[], rowBegin: int64[], aLocal: DnTensor[]) =
        this.Bracket (fun r ->
            let mutable t, a = tFull.[r], aLocal.[r]
            Native.check (Native.dn_shard_reduce_last_axis (group, r, op, &t, rowBegin.[r], &a)))

    member this.ArgReduceLastAxis (op: int, tFull: DnTensor[], rowBegin: int64[], aLocal: DnTensor[]) =
        this.Bracket (fun r ->
            let mutable t, a = tFull.[r], aLocal.[r]
            Native.check (Native.dn_shard_arg_reduce_last_axis (group, r, op, &t, rowBegin.[r], &a)))

    /// Max + ArgMax (op = 1) or Min + ArgMin (op = 0) in one pass over the slabs.
    member this.MinMaxArgLastAxis (op: int, tVal: DnTensor[], tIdx: DnTensor[], rowBegin: int64[], aLocal: DnTensor[]) =
        this.Bracket (fun r ->
            let mutable v, i, a = tVal.[r], tIdx.[r], aLocal.[r]
            Native.check (Native.dn_shard_minmax_arg_last_axis (group, r, op, &v, &i, rowBegin.[r], &a)))

    /// Reduction over the SHARDED axis (whole-tensor folds): kind 0 = fold, 1 = arg, 2 = find.
    member this.ReduceShardedAxis (kind: int, op: int, value: nativeint, t: DnTensor[], axisBegin: int64[], aLocal: DnTensor[]) =
        this.Bracket (fun r ->
            let mutable tt, a = t.[r], aLocal.[r]
            Native.check (Native.dn_shard_reduce_sharded_axis (group, r, kind, op, value, &tt, axisBegin.[r], &a)))

    /// countTrue of every rank's mask slab (blocking; _begin on every rank first, then _end: one thread drives all ranks).
    member this.CountTrue (maskLocal: DnTensor[]) =
        this.Bracket (fun r ->
            let mutable m = maskLocal.[r]
            Native.check (Native.dn_shard_count_true_begin (group, r, &m)))
        let counts = Array.zeroCreate<int64> world
        for r in 0 .. world - 1 do Native.check (Native.dn_shard_count_true_end (group, r, counts))
        counts

    /// Tensor.trueIdx of a bool tensor sharded along dim 0: returns the device pointer of each rank's full
    /// [nTrue, nDims] int64 result and nTrue.
    member this.TrueIndices (maskLocal: DnTensor[], dim0Begin: int64[]) =
        let counts = this.CountTrue maskLocal
        let total = Array.sum counts
        let nd = maskLocal.[0].NDims
        let ptrs = Array.init world (fun r -> this.HeapAlloc (r, max 1L (total * int64 nd * 8L)))
        this.Bracket (fun r ->
            let mutable t = Marshalling.desc ptrs.[r] (TensorLayout.newC [total; int64 nd]) DnDType.I64
            let mutable m = maskLocal.[r]
            let off = counts |> Seq.take r |> Seq.sum
            Native.check (Native.dn_shard_true_indices (group, r, &t, off, counts.[r], &m, dim0Begin.[r])))
        ptrs, total

    interface IDisposable with
        member this.Dispose () =
            if group <> 0n then
                Native.dn_shard_group_destroy group |> ignore
                group <- 0n
