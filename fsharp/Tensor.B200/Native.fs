namespace Tensor.B200

// P/Invoke binding of libdeepnet_b200.so (include/dn_tensor.h).
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no dotnet/fsharpc available). What IS checked here, on every test run
// (tests/test_abi.py): every `extern` below names a symbol libdeepnet_b200.so exports, every entry point the header
// declares has an `extern` here, and the argument / return ABI classes (pointer, int32, int64) of each pair agree;
// the struct layout (sequential, 152 bytes) is checked from C and from the ctypes mirror.
// Loading follows the reference's own mechanism for native libraries (Tensor/Tensor/NativeLib.fs:138-202,
// NativeLibName.Packaged "deepnet_b200" -> runtimes/linux-x64/native/libdeepnet_b200.so), exactly how
// Tensor/Tensor/Host/HostBLAS.fs:113-157 binds MKL.

open System
open System.Runtime.InteropServices

/// dn_dtype — order is ABI.
type DnDType =
    | F32 = 0 | F64 = 1 | I8 = 2 | U8 = 3 | I16 = 4 | U16 = 5
    | I32 = 6 | U32 = 7 | I64 = 8 | U64 = 9 | Bool = 10

/// dn_status and the exception raised for it.
type DnStatus =
    | Ok = 0 | InvalidArg = 1 | Unsupported = 2 | OutOfMemory = 3 | IndexOutOfRange = 4
    | Cuda = 5 | NoDevice = 6 | ShapeMismatch = 7 | SingularMatrix = 8

/// dn_tensor — replaces NativeTensor (Tensor/Tensor/Cuda/NativeTensor.fs:50-57). 152 bytes.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type DnTensor =
    val mutable Base:   nativeint
    val mutable Offset: int64
    val mutable NDims:  int32
    val mutable DType:  int32
    [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)>]
    val mutable Shape:  int64[]
    [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)>]
    val mutable Stride: int64[]

module Native =
    [<Literal>]
    let Lib = "deepnet_b200"

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_init(int device)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_device_count(int& count)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_device(int device)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_get_device(int& device)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_stream(nativeint stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_get_stream(nativeint& stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_release_stream(nativeint stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_sync()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_check_errors(int enabled)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_poll_index_error(int& hadError)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern nativeint dn_last_error()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern int64 dn_launch_count()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern nativeint dn_version()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_alloc(int64 nbytes, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_free(nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_free_deferred(nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_alloc_host(int64 nbytes, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_free_host(nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_host_register(nativeint ptr, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_host_unregister(nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memset_zero(nativeint ptr, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_h2d(nativeint dstDev, nativeint srcHost, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_d2h(nativeint dstHost, nativeint srcDev, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_d2d(nativeint dstDev, nativeint srcDev, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_d2h_async(nativeint dstHost, nativeint srcDev, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_transfer_h2d(DnTensor& devT, DnTensor& hostT)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_transfer_d2h(DnTensor& hostT, DnTensor& devT)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_event_create(nativeint& event)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_event_destroy(nativeint event)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_event_record(nativeint event)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_stream_wait_event(nativeint event)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_get_item(DnTensor& t, int64[] pos, nativeint value)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_item(DnTensor& t, int64[] pos, nativeint value)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fill_const(DnTensor& t, nativeint value)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fill_incrementing(DnTensor& t, nativeint start, nativeint incr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_copy(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_convert(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_unary(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_binary(int op, DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_compare(int op, DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_is_finite(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_if_then_else(DnTensor& t, DnTensor& cond, DnTensor& ifTrue, DnTensor& ifFalse)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_reduce_last_axis(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_arg_reduce_last_axis(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_find_last_axis(nativeint value, DnTensor& t, DnTensor& a)
    // index / mask tensors: array of pointers to pinned DnTensor structs, IntPtr.Zero = None / NoMask
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_gather(DnTensor& t, nativeint[] idxs, int nidxs, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_scatter(DnTensor& t, nativeint[] idxs, int nidxs, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_count_true(DnTensor& a, int64& count)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_masked_get(DnTensor& t, DnTensor& a, nativeint[] masks, int nmasks)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_masked_set(DnTensor& t, nativeint[] masks, int nmasks, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_true_indices(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_math_mode(int mode)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_get_math_mode(int& mode)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_vec_vec_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_mat_vec_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_mat_mat_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_batched_mat_mat_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_batched_invert(DnTensor& t, DnTensor& a)
    /// dn_fused_instr: kind (0 unary, 1 binary, 2 const), op, dst, a, b, imm — see include/dn_tensor.h
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type DnFusedInstr =
        val mutable Kind: int; val mutable Op: int; val mutable Dst: int; val mutable A: int; val mutable B: int
        val mutable Imm: double
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fused_elemwise(DnTensor& t, nativeint[] srcs, int nsrc, DnFusedInstr[] prog, int ninstr)

    // multi-GPU: leading-axis sharding (include/dn_tensor.h "Multi-GPU")
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_arg_reduce_combine(int op, DnTensor& t, DnTensor& vals, DnTensor& idxs)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_create(int world, int nlocal, int[] localRanks, int[] localDevices, int64 heapBytes, nativeint& group)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_handle(nativeint group, int rank, byte[] handle)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_connect(nativeint group, byte[] handles)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_destroy(nativeint group)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_set_stream(nativeint group, int rank, nativeint stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_sync(nativeint group, int rank)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_slab(int64 nrows, int rank, int world, int64& rowBegin, int64& count)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_heap_alloc(nativeint group, int rank, int64 nbytes, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_heap_reset(nativeint group, int rank)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_barrier(nativeint group, int rank)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_start(nativeint group)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_group_end(nativeint group)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_reduce_last_axis(nativeint group, int rank, int op, DnTensor& tFull, int64 rowBegin, DnTensor& aLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_arg_reduce_last_axis(nativeint group, int rank, int op, DnTensor& tFull, int64 rowBegin, DnTensor& aLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_find_last_axis(nativeint group, int rank, nativeint value, DnTensor& tFull, int64 rowBegin, DnTensor& aLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_minmax_arg_last_axis(nativeint group, int rank, int op, DnTensor& tValFull, DnTensor& tIdxFull, int64 rowBegin, DnTensor& aLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_reduce_sharded_axis(nativeint group, int rank, int kind, int op, nativeint value, DnTensor& t, int64 axisBegin, DnTensor& aLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_all_gather_rows(nativeint group, int rank, DnTensor& tFull, int64 rowBegin, int64 nrows)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_count_true(nativeint group, int rank, DnTensor& maskLocal, int64[] counts)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_count_true_begin(nativeint group, int rank, DnTensor& maskLocal)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_count_true_end(nativeint group, int rank, int64[] counts)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_true_indices(nativeint group, int rank, DnTensor& tFull, int64 rowOffset, int64 nrowsLocal, DnTensor& maskLocal, int64 dim0Begin)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_shard_masked_get(nativeint group, int rank, DnTensor& tFull, int64 elemOffset, int64 nelemsLocal, DnTensor& aLocal, DnTensor& maskLocal)

    /// Maps a non-OK status to the exception the reference raises in the same situation (SURVEY.md §8b).
    let check (st: DnStatus) =
        if st <> DnStatus.Ok then
            let msg = Marshal.PtrToStringAnsi (dn_last_error ())
            match st with
            | DnStatus.Unsupported      -> raise (NotSupportedException msg)
            | DnStatus.OutOfMemory      -> raise (OutOfMemoryException msg)   // OutOfCudaMemoryException in Tensor.Cuda
            | DnStatus.IndexOutOfRange  -> raise (IndexOutOfRangeException msg)
            | DnStatus.InvalidArg       -> raise (ArgumentException msg)
            | DnStatus.ShapeMismatch    -> raise (InvalidOperationException msg)
            | DnStatus.SingularMatrix   -> raise (SingularMatrixException msg)
            | _                         -> failwithf "CUDA error: %s" msg      // ManagedCuda.CudaException
