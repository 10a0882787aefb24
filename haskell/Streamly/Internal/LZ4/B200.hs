{-# LANGUAGE CPP #-}
-- |
-- Module      : Streamly.Internal.LZ4.B200
-- Portability : GHC
--
-- Batched replacements for 'compressChunksD' / 'decompressChunksRawD' of
-- "Streamly.Internal.LZ4" (src/Streamly/Internal/LZ4.hs:353-394, :539-567)
-- over libb200lz4.so (include/b200lz4.h).  Same stream types, same block
-- header layout, same acceleration parameter, same error texts; the only
-- behavioural difference is that up to 'batchArrays' arrays (or 'batchBytes'
-- bytes) are pulled from upstream before one @safe@ FFI call is made, and the
-- results are then yielded one by one in order.
--
-- STATUS: written against streamly-0.8's internal Array / StreamD API exactly
-- as the reference uses it, but NOT type-checked in this repository's build
-- image (no GHC there).  Every behaviour it relies on is exercised at the C
-- ABI: examples/ffi_replay.c makes exactly the call sequences below (same
-- argument marshalling: Int64 offsets, Int32 lengths, 16-byte staged gaps,
-- @[0, n]@ stream table or NULL for independent blocks) and runs in the GPU test
-- suite; tests/test_gpu_parity.py does the same through streamly_lz4_b200/api.py.
-- Drop this file next to src/Streamly/Internal/LZ4.hs, apply
-- haskell/patches/Config.hs.patch (it implements 'setBlockIndependence') and
-- re-export from there (INTEGRATION.md 2).
--
module Streamly.Internal.LZ4.B200
    ( compressChunksD
    , decompressChunksRawD
    , resizeChunksD
    , B200Config (..)
    , defaultB200Config
    )

where

import Control.Monad (forM, forM_, when)
import Control.Monad.IO.Class (MonadIO(..))
import Data.Int (Int32, Int64)
import Data.Word (Word8)
import Foreign.C (CInt(..), CSize(..), CString, peekCString)
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, pokeArray)
import Foreign.Marshal.Utils (copyBytes)
import Foreign.Ptr (Ptr, nullPtr, plusPtr)
import Foreign.Storable (peek, peekByteOff, poke)
import Fusion.Plugin.Types (Fuse (..))

import qualified Streamly.Internal.Data.Array.Foreign as Array
import qualified Streamly.Internal.Data.Array.Foreign.Type as Array
import qualified Streamly.Internal.Data.Array.Foreign.Mut.Type as MArray
import qualified Streamly.Internal.Data.Stream.StreamD as Stream

import Streamly.Internal.LZ4.Config

#define INLINE_NORMAL INLINE [1]
#define INLINE_LATE   INLINE [0]

--------------------------------------------------------------------------------
-- Foreign (include/b200lz4.h)
--------------------------------------------------------------------------------

data C_Ctx
data C_CStream
data C_DStream

foreign import ccall unsafe "b200lz4.h b200lz4_compress_bound"
    c_compressBound :: CInt -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_ctx_create"
    c_ctxCreate :: CInt -> Ptr (Ptr C_Ctx) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_ctx_destroy"
    c_ctxDestroy :: Ptr C_Ctx -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_host_alloc"
    c_hostAlloc :: CSize -> IO (Ptr Word8)

foreign import ccall safe "b200lz4.h b200lz4_host_free"
    c_hostFree :: Ptr Word8 -> IO ()

-- The error text is kept in the ctx, not in thread-local storage: an unbound Haskell thread may run the @safe@ batch call
-- and this call on different OS threads.
foreign import ccall unsafe "b200lz4.h b200lz4_ctx_last_error"
    c_ctxLastError :: Ptr C_Ctx -> IO CString

foreign import ccall unsafe "b200lz4.h b200lz4_last_error"
    c_lastError :: IO CString

-- parallel gather of the batch's arrays into the page-locked source buffer (host threads inside the library)
foreign import ccall safe "b200lz4.h b200lz4_gather_host"
    c_gatherHost :: Ptr Word8 -> Ptr (Ptr Word8) -> Ptr Int64 -> Ptr Int32 -> CInt -> CInt -> IO CInt

-- resizeChunksD's header walk over one contiguous range
foreign import ccall unsafe "b200lz4.h b200lz4_reframe"
    c_reframe :: Ptr Word8 -> Int64 -> CInt -> CInt -> Ptr Int64 -> Ptr Int32 -> Int64
              -> Ptr Int64 -> Ptr Int64 -> Ptr CInt -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_cstream_create"
    c_cstreamCreate :: Ptr C_Ctx -> Ptr (Ptr C_CStream) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_cstream_free"
    c_cstreamFree :: Ptr C_CStream -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_dstream_create"
    c_dstreamCreate :: Ptr C_Ctx -> Ptr (Ptr C_DStream) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_dstream_free"
    c_dstreamFree :: Ptr C_DStream -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_compress_batch"
    c_compressBatch
        :: Ptr C_Ctx
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32 -> CInt
        -> Ptr Int32 -> CInt -> Ptr (Ptr C_CStream)
        -> CInt -> CInt
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32
        -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_decompress_batch"
    c_decompressBatch
        :: Ptr C_Ctx
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32 -> CInt
        -> Ptr Int32 -> CInt -> Ptr (Ptr C_DStream)
        -> CInt -> CInt
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32
        -> IO CInt

lz4_MAX_INPUT_SIZE :: Int
lz4_MAX_INPUT_SIZE = 0x7E000000           -- B200LZ4_MAX_INPUT_SIZE

--------------------------------------------------------------------------------
-- Batching parameters
--------------------------------------------------------------------------------

data B200Config = B200Config
    { device :: Int            -- ^ CUDA device ordinal
    , batchArrays :: Int       -- ^ at most this many arrays per FFI call
    , batchBytes :: Int        -- ^ ... or this many input bytes
    }

defaultB200Config :: B200Config
defaultB200Config = B200Config 0 4096 (256 * 1024 * 1024)

-- | One library context plus two growable page-locked staging buffers.
data Session = Session
    { sCtx :: Ptr C_Ctx
    , sSrc :: Ptr Word8, sSrcCap :: Int
    , sDst :: Ptr Word8, sDstCap :: Int
    }

lastError :: IO String
lastError = c_lastError >>= peekCString

ctxError :: Session -> IO String
ctxError s = c_ctxLastError (sCtx s) >>= peekCString

newSession :: B200Config -> IO Session
newSession conf = alloca $ \pp -> do
    rc <- c_ctxCreate (fromIntegral (device conf)) pp
    when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_ctx_create failed: " ++ e)
    ctx <- peek pp
    let cap = batchBytes conf + 16 * (batchArrays conf + 1)
    src <- c_hostAlloc (fromIntegral cap)
    return $ Session ctx src cap nullPtr 0

freeSession :: Session -> IO ()
freeSession s = do
    c_hostFree (sSrc s)
    when (sDst s /= nullPtr) $ c_hostFree (sDst s)
    c_ctxDestroy (sCtx s)

ensureDst :: Session -> Int -> IO Session
ensureDst s need
    | need <= sDstCap s = return s
    | otherwise = do
        when (sDst s /= nullPtr) $ c_hostFree (sDst s)
        p <- c_hostAlloc (fromIntegral (need + need `div` 4))
        return s { sDst = p, sDstCap = need + need `div` 4 }

ensureSrc :: Session -> Int -> IO Session
ensureSrc s need
    | need <= sSrcCap s = return s
    | otherwise = do
        c_hostFree (sSrc s)
        p <- c_hostAlloc (fromIntegral (need + need `div` 4))
        return s { sSrc = p, sSrcCap = need + need `div` 4 }

stagedSize :: [Array.Array Word8] -> Int
stagedSize = sum . map (\a -> align16 (Array.byteLength a + 16))

align16 :: Int -> Int
align16 n = (n + 15) `div` 16 * 16

-- | Copy the arrays into the pinned source buffer at 16-byte aligned offsets
-- with a 16-byte gap (separate Haskell arrays are never adjacent) and return
-- (offsets, lengths).  The copy itself runs on the library's host threads
-- (b200lz4_gather_host); the arrays are pinned (streamly allocates them so),
-- so their addresses stay valid for the duration of the call.
stage :: Session -> [Array.Array Word8] -> IO ([Int64], [Int32])
stage s arrs = do
    let lens = map Array.byteLength arrs
        offs = scanl (\at n -> at + align16 (n + 16)) 0 lens
        n = length arrs
    withPtrs arrs [] $ \ptrs ->
      allocaArray n $ \pPtr -> allocaArray n $ \pOff -> allocaArray n $ \pLen -> do
        pokeArray pPtr ptrs
        pokeArray pOff (map fromIntegral (take n offs))
        pokeArray pLen (map fromIntegral lens)
        rc <- c_gatherHost (sSrc s) pPtr pOff pLen (fromIntegral n) 0
        when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_gather_host failed: " ++ e)
    return (map fromIntegral (take n offs), map fromIntegral lens)
  where
    withPtrs [] acc k = k (reverse acc)
    withPtrs (a:as) acc k =
        Array.asPtrUnsafe (Array.unsafeCast a) $ \p -> withPtrs as ((p :: Ptr Word8) : acc) k

-- | Slice a fresh, exactly sized array out of the pinned destination buffer.
sliceOut :: Ptr Word8 -> Int -> Int -> IO (Array.Array Word8)
sliceOut base off len = do
    (MArray.Array cont b0 b dmax) <- MArray.newArray len
    copyBytes b (base `plusPtr` off) len
    return $ Array.unsafeFreeze (MArray.Array cont b0 (b `plusPtr` len) dmax)

--------------------------------------------------------------------------------
-- One batch through the library
--------------------------------------------------------------------------------

maxBlockSizeOf :: BlockConfig -> Int
maxBlockSizeOf cfg =
    case blockSize cfg of
        BlockHasSize -> lz4_MAX_INPUT_SIZE
        BlockMax64KB -> 64 * 1024
        BlockMax256KB -> 256 * 1024
        BlockMax1MB -> 1024 * 1024
        BlockMax4MB -> 4 * 1024 * 1024

-- | 'compressChunk' (src/Streamly/Internal/LZ4.hs:226-281) for a whole batch.
compressBatch
    :: BlockConfig -> Int -> Session -> Ptr C_CStream -> [Array.Array Word8]
    -> IO (Session, [Array.Array Word8])
compressBatch cfg speed s0 strm arrs = do
    let n = length arrs
        meta = metaSize cfg
    forM_ arrs $ \a -> do
        let len = Array.byteLength a
        when (len >= 2 * 1024 * 1024 * 1024)
            $ error "compressChunksD: Array element > 2 GB encountered"
        when (len > maxBlockSizeOf cfg)
            $ error $ "compressChunk: Source array length " ++ show len
                ++ " exceeds the maximum block size of " ++ show (maxBlockSizeOf cfg)
    bounds <- forM arrs $ \a -> do
        b <- c_compressBound (fromIntegral (Array.byteLength a))
        when (b <= 0) $ error "compressChunk: compressed length <= 0."
        return (fromIntegral b + meta)
    s <- ensureDst s0 (sum bounds) >>= \s' -> ensureSrc s' (stagedSize arrs)
    (offs, lens) <- stage s arrs
    let srcBytes = case (offs, lens) of
            ([], _) -> 0
            _ -> fromIntegral (last offs) + fromIntegral (last lens)
    allocaArray n $ \pOff -> allocaArray n $ \pLen ->
      allocaArray (n + 1) $ \pDstOff -> allocaArray n $ \pOutLen ->
      allocaArray 2 $ \pFirst -> alloca $ \pStrm -> do
        pokeArray pOff offs
        pokeArray pLen lens
        pokeArray pFirst [0, fromIntegral n]
        poke pStrm strm
        -- linked mode: all blocks of the batch go through the stream's one state, in order
        -- (one LZ4_stream_t per Haskell stream, :367-376); independent: no stream table
        let (pF, nS, pS) = if strm == nullPtr then (nullPtr, 0, nullPtr) else (pFirst, 1, pStrm)
        rc <- c_compressBatch (sCtx s) (sSrc s) srcBytes pOff pLen (fromIntegral n)
                  pF nS pS (fromIntegral speed) (fromIntegral meta)
                  (sDst s) (fromIntegral (sDstCap s)) pDstOff pOutLen
        when (rc /= 0) $ ctxError s >>= \e ->
            error ("compressChunk: c_compressFastContinue failed. " ++ e)
        dstOff <- peekArray (n + 1) pDstOff
        outs <- forM (zip dstOff (tail dstOff)) $ \(a, b) ->
            sliceOut (sDst s) (fromIntegral a) (fromIntegral (b - a))
        return (s, outs)

-- | 'decompressChunk' (src/Streamly/Internal/LZ4.hs:290-336) for a whole batch of framed arrays.
decompressBatch
    :: BlockConfig -> Session -> Ptr C_DStream -> [Array.Array Word8]
    -> IO (Session, [Array.Array Word8])
decompressBatch cfg s0 strm arrs = do
    let n = length arrs
        meta = metaSize cfg
        maxBlock = case blockSize cfg of
            BlockHasSize -> 0
            _ -> maxBlockSizeOf cfg
    caps <- forM arrs $ \a ->
        if meta == 8 && Array.byteLength a >= 8
        then Array.asPtrUnsafe (Array.unsafeCast a) $ \p -> do
                 u <- peekByteOff (p :: Ptr Word8) 4 :: IO Int32     -- uncompLen LE32 (little-endian host)
                 return (max 0 (fromIntegral u))
        else return maxBlock
    s <- ensureDst s0 (sum caps + 64) >>= \s' -> ensureSrc s' (stagedSize arrs)
    (offs, lens) <- stage s arrs
    let srcBytes = case (offs, lens) of
            ([], _) -> 0
            _ -> fromIntegral (last offs) + fromIntegral (last lens)
    allocaArray n $ \pOff -> allocaArray n $ \pLen ->
      allocaArray (n + 1) $ \pDstOff -> allocaArray n $ \pOutLen ->
      allocaArray 2 $ \pFirst -> alloca $ \pStrm -> do
        pokeArray pOff offs
        pokeArray pLen lens
        pokeArray pFirst [0, fromIntegral n]
        poke pStrm strm
        let (pF, nS, pS) = if strm == nullPtr then (nullPtr, 0, nullPtr) else (pFirst, 1, pStrm)
        rc <- c_decompressBatch (sCtx s) (sSrc s) srcBytes pOff pLen (fromIntegral n)
                  pF nS pS (fromIntegral meta) (fromIntegral maxBlock)
                  (sDst s) (fromIntegral (sDstCap s)) pDstOff pOutLen
        when (rc /= 0) $ ctxError s >>= \e ->
            error ("decompressChunk: c_decompressSafeContinue failed. " ++ e)
        dstOff <- peekArray n pDstOff
        outLen <- peekArray n pOutLen
        outs <- forM (zip dstOff outLen) $ \(a, l) ->
            sliceOut (sDst s) (fromIntegral a) (fromIntegral l)
        return (s, outs)

--------------------------------------------------------------------------------
-- Stream combinators
--------------------------------------------------------------------------------

{-# ANN type BatchState Fuse #-}
data BatchState st ses
    = BInit st
    | BFill st ses [Array.Array Word8] Int Int     -- pending arrays (reversed), count, bytes
    | BDrain st ses [Array.Array Word8] Bool       -- ready outputs; True = upstream finished
    | BDone ses

-- | Shared driver: pull up to a batch, run it, yield the results in order.
{-# INLINE_NORMAL batchedD #-}
batchedD ::
       MonadIO m
    => B200Config
    -> IO ses                                              -- ^ acquire (context, stream state)
    -> (ses -> IO ())                                      -- ^ release
    -> (ses -> [Array.Array Word8] -> IO (ses, [Array.Array Word8]))
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
batchedD conf acquire release run (Stream.Stream step0 state0) =
    Stream.Stream step (BInit state0)

    where

    flush st ses pending finished = do
        (ses1, outs) <- liftIO $ run ses (reverse pending)
        return $ Stream.Skip $ BDrain st ses1 outs finished

    {-# INLINE_LATE step #-}
    step _ (BInit st) = do
        ses <- liftIO acquire
        return $ Stream.Skip $ BFill st ses [] 0 0
    step gst (BFill st ses pending cnt bytes)
        | cnt >= batchArrays conf || bytes >= batchBytes conf = flush st ses pending False
        | otherwise = do
            r <- step0 gst st
            case r of
                Stream.Yield arr st1 ->
                    return $ Stream.Skip
                        $ BFill st1 ses (arr : pending) (cnt + 1) (bytes + Array.byteLength arr)
                Stream.Skip st1 -> return $ Stream.Skip $ BFill st1 ses pending cnt bytes
                Stream.Stop ->
                    if null pending
                    then return $ Stream.Skip $ BDone ses
                    else flush st ses pending True
    step _ (BDrain st ses (a : as) fin) = return $ Stream.Yield a (BDrain st ses as fin)
    step _ (BDrain st ses [] False) = return $ Stream.Skip $ BFill st ses [] 0 0
    step _ (BDrain _ ses [] True) = return $ Stream.Skip $ BDone ses
    step _ (BDone ses) = liftIO (release ses) >> return Stream.Stop

-- | Drop-in for 'Streamly.Internal.LZ4.compressChunksD' (:353-394).  Linked
-- blocks by default (one device-resident stream state per Haskell stream);
-- with @setBlockIndependence True@ (implemented by
-- haskell/patches/Config.hs.patch; a stub in the reference, Config.hs:142-146)
-- every array is compressed with a fresh state: no stream handle is created and
-- the batch call gets a NULL stream table, which is the mode in which all blocks
-- of a batch run in parallel.
{-# INLINE_NORMAL compressChunksD #-}
compressChunksD ::
       MonadIO m
    => B200Config
    -> BlockConfig
    -> Int
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
compressChunksD conf cfg speed0 =
    batchedD conf acquire release run

    where

    speed = max speed0 0                                        -- :364
    acquire = do
        s <- newSession conf
        strm <- if blockIndependent cfg
            then return nullPtr
            else alloca $ \pp -> do
                rc <- c_cstreamCreate (sCtx s) pp
                when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_cstream_create failed: " ++ e)
                peek pp
        return (s, strm)
    release (s, strm) = when (strm /= nullPtr) (c_cstreamFree strm) >> freeSession s      -- :393-394
    run (s, strm) arrs = do
        (s1, outs) <- compressBatch cfg speed s strm arrs
        return ((s1, strm), outs)

-- | Drop-in for 'Streamly.Internal.LZ4.decompressChunksRawD' (:539-567): every
-- input array is exactly one framed block (what 'resizeChunksD' yields).
{-# INLINE_NORMAL decompressChunksRawD #-}
decompressChunksRawD ::
       MonadIO m
    => B200Config
    -> BlockConfig
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
decompressChunksRawD conf cfg =
    batchedD conf acquire release run

    where

    acquire = do
        s <- newSession conf
        strm <- if blockIndependent cfg
            then return nullPtr
            else alloca $ \pp -> do
                rc <- c_dstreamCreate (sCtx s) pp
                when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_dstream_create failed: " ++ e)
                peek pp
        return (s, strm)
    release (s, strm) = when (strm /= nullPtr) (c_dstreamFree strm) >> freeSession s
    run (s, strm) arrs = do
        (s1, outs) <- decompressBatch cfg s strm arrs
        return ((s1, strm), outs)

--------------------------------------------------------------------------------
-- Re-framing
--------------------------------------------------------------------------------

{-# ANN type ResizeState Fuse #-}
data ResizeState st
    = RInit st
    | RHave st (Array.Array Word8)                 -- unconsumed bytes
    | RYield st (Array.Array Word8) [Array.Array Word8] Bool   -- rest, ready blocks, end mark seen
    | RDone

-- | Drop-in for 'Streamly.Internal.LZ4.resizeChunksD' (:432-523): the header
-- walk of one accumulated range is one call of b200lz4_reframe (which consumes
-- as many complete [header][block] arrays as the range holds and reports the
-- end mark); the state machine around it -- accumulate while a block is
-- incomplete, stop at the end mark, "Incomplete block" / "No end mark found"
-- at the end of the input -- is the reference's.
{-# INLINE_NORMAL resizeChunksD #-}
resizeChunksD ::
       MonadIO m
    => BlockConfig
    -> FrameConfig
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
resizeChunksD cfg frameCfg (Stream.Stream step0 state0) =
    Stream.Stream step (RInit state0)

    where

    meta = metaSize cfg
    endMark = hasEndMark frameCfg
    maxBlocks = 4096 :: Int

    -- all complete blocks at the front of `arr`: (blocks, rest, end mark seen)
    walk arr = liftIO $ do
        let len = Array.byteLength arr
        Array.asPtrUnsafe (Array.unsafeCast arr) $ \p ->
          allocaArray maxBlocks $ \pOff -> allocaArray maxBlocks $ \pLen ->
          alloca $ \pN -> alloca $ \pUsed -> alloca $ \pEnd -> do
            rc <- c_reframe (p :: Ptr Word8) (fromIntegral len) (fromIntegral meta)
                      (if endMark then 1 else 0) pOff pLen (fromIntegral maxBlocks) pN pUsed pEnd
            when (rc /= 0) $ lastError >>= \e -> error ("resizeChunksD: " ++ e)
            n <- fromIntegral <$> peek pN
            used <- fromIntegral <$> peek pUsed
            ended <- peek pEnd
            offs <- peekArray n pOff
            lens <- peekArray n pLen
            let blocks = [ Array.getSliceUnsafe (fromIntegral o) (fromIntegral l) arr | (o, l) <- zip offs lens ]
                rest = Array.getSliceUnsafe used (len - used) arr
            return (blocks, rest, ended /= 0)

    {-# INLINE_LATE step #-}
    step gst (RInit st) = do
        r <- step0 gst st
        case r of
            Stream.Yield arr st1 -> return $ Stream.Skip $ RHave st1 arr
            Stream.Skip st1 -> return $ Stream.Skip $ RInit st1
            Stream.Stop ->
                if endMark
                then error "resizeChunksD: No end mark found"          -- :493-496
                else return Stream.Stop
    step gst (RHave st arr) = do
        (blocks, rest, ended) <- walk arr
        if null blocks && not ended
        then do                                                      -- RAccumulate, :498-505
            r <- step0 gst st
            case r of
                Stream.Yield more st1 -> do
                    arr1 <- liftIO $ Array.spliceTwo arr more
                    return $ Stream.Skip $ RHave st1 arr1
                Stream.Skip st1 -> return $ Stream.Skip $ RHave st1 arr
                Stream.Stop ->
                    if Array.byteLength arr == 0 && not endMark
                    then return Stream.Stop
                    else error "resizeChunksD: Incomplete block"     -- :505
        else return $ Stream.Skip $ RYield st rest blocks ended
    step _ (RYield st rest (b : bs) ended) = return $ Stream.Yield b (RYield st rest bs ended)
    step _ (RYield st rest [] ended)
        | ended = return $ Stream.Skip RDone                         -- RFooter, :506-522: the stream stops at the end mark
        | Array.byteLength rest == 0 = return $ Stream.Skip $ RInit st
        | otherwise = return $ Stream.Skip $ RHave st rest
    step _ RDone = return Stream.Stop
